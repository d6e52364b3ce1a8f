#!/bin/bash
# One gpurun call: diagnostics, GPU tests, smoke, short bench. Everything logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== debug"; timeout 900 python tools/gpu_debug.py 2>&1 | tee gpurun_out/debug.log | tail -60
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -q -x --no-header 2>&1 | tee gpurun_out/pytest.log | tail -25
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log | tail -5
echo "=== bench"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tee gpurun_out/bench.log | tail -3
