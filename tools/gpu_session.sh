#!/bin/bash
# One gpurun call: diagnostics, GPU tests, smoke, short bench. Everything logged under gpurun_out/.
mkdir -p gpurun_out
echo "=== debug"; timeout 900 python tools/gpu_debug.py 2>&1 | tee gpurun_out/debug.log | tail -70
echo "=== pytest"; timeout 1200 python -m pytest tests -m gpu -q --no-header 2>&1 | tee gpurun_out/pytest.log | tail -25
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log | tail -5
echo "=== bench v2"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/bench.log | tail -3
echo "=== bench v1"; timeout 900 python bench.py --steps 5 --warmup 3 --tc-variant 1 --no-e2e --no-cpu 2>&1 | tee gpurun_out/bench_v1.log | tail -3 | cut -c1-600
