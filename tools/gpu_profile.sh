#!/bin/bash
# ncu evidence for the dominant kernels (one GPU). Numbers printed under ncu are NOT bench values.
mkdir -p gpurun_out
R=${1:-r01c}
echo "=== launch list (same command as the bench line, short)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_${R}.log 2>&1
tail -1 gpurun_out/launches_${R}.log | cut -c1-200
echo "=== full capture: tensor-core kernel (CTA-pair)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:l2_top2_tc2 -s 4 -c 1 -o gpurun_out/prof_tc2_${R} -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_tc2_${R}.log 2>&1
tail -1 gpurun_out/prof_tc2_${R}.log | cut -c1-200
echo "=== bench (final line of the round)"
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/bench_${R}.log | tail -1 | cut -c1-300
echo "=== reference arm"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tee gpurun_out/bench_ref_${R}.log | tail -1 | cut -c1-600
