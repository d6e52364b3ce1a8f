#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tee gpurun_out/pytest.log | tail -15
echo "=== upload experiment"; timeout 900 python tools/gpu_upload_exp.py 2>&1 | tee gpurun_out/upload_exp.log | tail -12
