#!/bin/bash
# BASELINE config 5: feature-count sweep, kernel-boundary pairs/s vs the tensor roofline next to the CPU reference.
mkdir -p gpurun_out
: > gpurun_out/sweep.jsonl
for spec in "1024 200" "2048 160" "4096 120" "8192 100" "16384 48" "32768 24"; do
  set -- $spec
  timeout 900 python bench.py --features $1 --images $2 --steps 3 --warmup 3 --cpu-seconds 4 >> gpurun_out/sweep.jsonl 2>gpurun_out/sweep_err.log || echo "{\"failed\": \"$spec\"}" >> gpurun_out/sweep.jsonl
done
python - <<'PY'
import json
print("| features/img | images | pairs | value pairs/s | kernel TFLOP/s | frac of sustained peak | e2e pairs/s | CPU reference pairs/s (threads) |")
print("|---:|---:|---:|---:|---:|---:|---:|---:|")
for l in open("gpurun_out/sweep.jsonl"):
    d = json.loads(l)
    if "failed" in d: print("| failed", d["failed"], "|"); continue
    w = d["config"]["workload"].split()
    print(f"| {w[4]} | {w[0]} | {int(d['config']['pairs_per_gpu'])} | {d['value']:.0f} | {d['roofline']['achieved']:.0f} | {d['roofline']['frac']:.3f} | {d['e2e']['value']:.0f} | {d['cpu_baseline']['value']:.2f} ({d['cpu_baseline']['cores']}) |")
PY
