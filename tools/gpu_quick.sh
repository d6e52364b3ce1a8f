#!/bin/bash
# Quick A/B: parity of both tensor-core variants + kernel-boundary bench lines.
mkdir -p gpurun_out
echo "=== pytest (tensor-core subset)"; timeout 900 python -m pytest tests -m gpu -q --no-header -k "tensorcore or full_size or golden" 2>&1 | tee gpurun_out/pytest_quick.log | tail -6
for v in 4 2; do
  echo "=== bench variant $v"; timeout 600 python bench.py --steps 5 --warmup 3 --tc-variant $v --no-e2e --no-cpu 2>&1 | tee gpurun_out/bench_quick_v$v.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print({k: d[k] for k in ('value','ms_per_step')}, d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['clocks'])
"
done
