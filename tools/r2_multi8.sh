#!/bin/bash
N=$1; tag=$2
mkdir -p gpurun_out
line() { python tools/bench_line.py "$1"; }
runN() { name=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
           bench.py --gpus $N --steps 100 --warmup 3 --skip-cpu "$@" > gpurun_out/r2_${tag}_n${N}_$name.json 2> gpurun_out/r2_${tag}_n${N}_$name.err; \
           line n${N}_$name < gpurun_out/r2_${tag}_n${N}_$name.json; grep -i "error\|Traceback" gpurun_out/r2_${tag}_n${N}_$name.err | head -3; \
           python -c "import json,sys; d=json.loads([l for l in open('gpurun_out/r2_${tag}_n${N}_$name.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['step_breakdown_ms']['traverse'], d.get('parity',{}).get('ok'), d.get('structural'))"; }
runN nccl
runN peer --exchange peer --skip-variants
runN structural --structural --steps 20
