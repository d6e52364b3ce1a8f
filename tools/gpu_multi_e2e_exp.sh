#!/bin/bash
# e2e of the 4-rank bench under different host-side settings (staging copies in flight, finishing threads per rank)
mkdir -p gpurun_out
run() { echo "--- $1"; env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'e2e', round(d['e2e']['value']))"; }
run "B200M_UP_LAG=12" 29601
run "B200M_UP_LAG=4" 29602
run "B200M_UP_LAG=6 B200M_HOST_THREADS=16" 29603
run "B200M_UP_LAG=12 B200M_HOST_THREADS=12" 29604
