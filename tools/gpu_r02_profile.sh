#!/bin/bash
# Round-2 evidence on ONE GPU: launch list + ncu --set full of the default kernel and of the real-valued filter kernel, the bench
# lines (integer / real-valued / Hamming), the reference arm, the feature sweep.  Numbers printed under ncu are NOT bench values.
export B200M_NO_BUILD=1
R=${1:-r02}
O=gpurun_out/$R
mkdir -p $O
echo "=== launch list (same command as the bench line, short)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $O/launches.log 2>&1
echo "=== ncu --set full: default tensor-core kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:l2_top2_tc2 -s 4 -c 1 -o $O/prof_tc2 -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/prof_tc2.log 2>&1
ncu -i $O/prof_tc2.ncu-rep --page raw --csv > $O/prof_tc2_raw.csv 2>/dev/null
echo "=== ncu --set full: real-valued filter kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:l2_top2_tc2 -s 4 -c 1 -o $O/prof_real -f \
    python bench.py --data real --steps 1 --warmup 1 --no-e2e --no-cpu > $O/prof_real.log 2>&1
ncu -i $O/prof_real.ncu-rep --page raw --csv > $O/prof_real_raw.csv 2>/dev/null
rm -f $O/prof_real.ncu-rep      # keep the merged output small: the raw csv is what profiles/ cites
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_real.csv \
    python bench.py --data real --steps 2 --warmup 1 --no-cpu --no-e2e > $O/launches_real.log 2>&1
echo "=== bench lines"
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.json
timeout 600 python bench.py --data real --steps 5 --warmup 3 --cpu-seconds 8 > $O/bench_real.json 2> $O/bench_real.err; tail -c 300 $O/bench_real.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err; tail -c 300 $O/bench_reference_arm.json
timeout 600 python bench.py --dtype bin --features 16384 --images 40 --steps 3 --warmup 3 --cpu-seconds 8 > $O/bench_hamming_40img.json 2> $O/bench_ham.err; tail -c 300 $O/bench_hamming_40img.json
ls -la $O | head -30
