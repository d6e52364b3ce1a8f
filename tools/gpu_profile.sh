#!/bin/bash
# ncu evidence for the dominant kernels (one GPU). Numbers printed under ncu are NOT bench values.
mkdir -p gpurun_out
R=${1:-r01}
echo "=== launch list (same command as the bench line, short)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_${R}.log 2>&1
tail -2 gpurun_out/launches_${R}.log | cut -c1-300
echo "=== full capture: tensor-core kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:l2_top2_tc -s 4 -c 1 -o gpurun_out/prof_tc_${R} -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_tc_${R}.log 2>&1
tail -1 gpurun_out/prof_tc_${R}.log | cut -c1-200
echo "=== full capture: hamming kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hamming_top2 -s 1 -c 1 -o gpurun_out/prof_ham_${R} -f \
    python bench.py --dtype bin --features 16384 --images 16 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_ham_${R}.log 2>&1
tail -1 gpurun_out/prof_ham_${R}.log | cut -c1-200
echo "=== hamming bench (config 4 shape, reduced image count)"
timeout 600 python bench.py --dtype bin --features 16384 --images 40 --steps 3 --warmup 3 --cpu-seconds 8 2>&1 | tee gpurun_out/bench_ham_${R}.log | tail -1 | cut -c1-1500
ls -la gpurun_out
