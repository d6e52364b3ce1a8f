#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -3
for spec in "1024 200" "2048 160" "4096 120" "8192 100"; do
  set -- $spec
  timeout 600 python bench.py --features $1 --images $2 --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['achieved']), round(d['roofline']['frac'],3))"
done
