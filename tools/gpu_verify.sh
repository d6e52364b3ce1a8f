#!/bin/bash
# One gpurun call that re-checks the round's state on a B200: GPU tests, smoke, the bench line and the reference arm.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tee gpurun_out/pytest.log | tail -6
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log | tail -3
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/bench.log | tail -1 | cut -c1-400
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tee gpurun_out/bench_ref.log | tail -1 | cut -c1-300
