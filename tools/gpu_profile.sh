#!/bin/bash
# ncu evidence for the dominant kernels + the round's bench lines (one GPU). Numbers printed under ncu are NOT bench values.
mkdir -p gpurun_out
R=${1:-r01d}
echo "=== launch list (same command as the bench line, short)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_${R}.log 2>&1
tail -1 gpurun_out/launches_${R}.log | cut -c1-200
echo "=== full capture: tensor-core kernel (CTA-pair)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:l2_top2_tc2 -s 4 -c 1 -o gpurun_out/prof_tc2_${R} -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_tc2_${R}.log 2>&1
tail -1 gpurun_out/prof_tc2_${R}.log | cut -c1-200
echo "=== full capture: Hamming kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hamming_top2 -s 1 -c 1 -o gpurun_out/prof_ham_${R} -f \
    python bench.py --dtype bin --features 16384 --images 16 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/prof_ham_${R}.log 2>&1
tail -1 gpurun_out/prof_ham_${R}.log | cut -c1-200
echo "=== bench (final line of the round)"
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/bench_${R}.log | tail -1 | cut -c1-300
echo "=== reference arm"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tee gpurun_out/bench_ref_${R}.log | tail -1 | cut -c1-600
echo "=== Hamming bench line (config 4 shape, 40 images)"
timeout 900 python bench.py --dtype bin --features 16384 --images 40 --steps 3 --warmup 3 --cpu-seconds 8 2>&1 | tee gpurun_out/bench_hamming_${R}.log | tail -1 | cut -c1-300
echo "=== raw csv exports of the captures"
for k in tc2 ham; do
  ncu -i gpurun_out/prof_${k}_${R}.ncu-rep --page raw --csv > gpurun_out/prof_${k}_${R}_raw.csv 2>/dev/null
  wc -c gpurun_out/prof_${k}_${R}_raw.csv
done
