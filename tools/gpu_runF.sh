#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tee gpurun_out/pytest.log | tail -15
echo "=== voctree bench"; timeout 900 python tools/gpu_voctree_bench.py 2>&1 | tee gpurun_out/voctree_bench.log | tail -3
