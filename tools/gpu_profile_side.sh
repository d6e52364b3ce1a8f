#!/bin/bash
# ncu --set full captures of the kernels either side of the path (quantiser, posting-list scorer, guided matching).
mkdir -p gpurun_out
R=${1:-r01d}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quantize_kernel -s 2 -c 1 -o gpurun_out/prof_quant_${R} -f \
    python tools/gpu_voctree_bench.py --images 24 --levels 5 --no-cpu > gpurun_out/prof_quant_${R}.log 2>&1; tail -1 gpurun_out/prof_quant_${R}.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_postings -c 1 -o gpurun_out/prof_score_${R} -f \
    python tools/gpu_voctree_bench.py --images 300 --levels 4 --no-cpu > gpurun_out/prof_score_${R}.log 2>&1; tail -1 gpurun_out/prof_score_${R}.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:guided_top2 -s 1 -c 1 -o gpurun_out/prof_guided_${R} -f \
    python tools/gpu_guided_bench.py --no-cpu > gpurun_out/prof_guided_${R}.log 2>&1; tail -1 gpurun_out/prof_guided_${R}.log | cut -c1-200
for k in quant score guided; do ncu -i gpurun_out/prof_${k}_${R}.ncu-rep --page raw --csv > gpurun_out/prof_${k}_${R}_raw.csv 2>/dev/null; wc -c gpurun_out/prof_${k}_${R}_raw.csv; done
