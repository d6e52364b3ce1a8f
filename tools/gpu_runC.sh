#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tee gpurun_out/pytest.log | tail -15
echo "=== e2e breakdown"; B200M_TIMING=1 timeout 600 python tools/gpu_e2e_breakdown.py 2>&1 | tee gpurun_out/e2e_breakdown.log | grep -v "^\[b200m\] match_pairs:" | tail -22
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/bench.log | tail -1 | cut -c1-2200
echo "=== bench config 3 (1000 images, voctree list) on one GPU"; timeout 1200 python bench.py --steps 3 --warmup 3 --images 1000 --pairs voctree --cpu-seconds 5 2>&1 | tee gpurun_out/bench_cfg3.log | tail -1 | cut -c1-1500
