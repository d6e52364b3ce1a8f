#!/bin/bash
# bench.py under torchrun on N GPUs of one box (the driver's launch line), weak scaling + optionally the fixed lists (configs 2 / 3).
export B200M_NO_BUILD=1
N=${1:-2}; R=${2:-r02}; EXTRA=${3:-}
O=gpurun_out/$R
mkdir -p $O
nvidia-smi topo -m > $O/topo_${N}gpu.txt 2>&1
cat /sys/fs/cgroup/cpu.max > $O/cpu_max_${N}gpu.txt 2>&1
run() {  # name, args...
  local name=$1; shift
  B200M_TIMING=${TIMING:-0} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > $O/${name}_${N}gpu.json 2> $O/${name}_${N}gpu.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/${name}_${N}gpu.json").read().strip().splitlines()[-1])
    print("$name N=$N value %.0f pairs/s, e2e %.0f pairs/s, ms/step %.1f, views/gpu max %s, host %s, numa %s" % (d["value"], d["e2e"]["value"] or 0, d["ms_per_step"], d["config"]["max_views_on_a_gpu"], d["config"]["host"], d["config"]["numa"]))
except Exception as e:
    print("$name N=$N FAILED", e); print(open("$O/${name}_${N}gpu.err").read()[-1500:])
PY
}
run bench --steps 5 --warmup 3
if [ "$EXTRA" = "rows" ]; then run bench_rows --steps 5 --warmup 3 --sharding rows; fi
if [ "$EXTRA" = "config2" ]; then run bench_config2 --config 2 --steps 2 --warmup 1 --no-cpu; fi
if [ "$EXTRA" = "config3" ]; then run bench_config3 --config 3 --steps 1 --warmup 1 --no-cpu; fi
