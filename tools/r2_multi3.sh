#!/bin/bash
N=$1; tag=$2
mkdir -p gpurun_out
line() { python tools/bench_line.py "$1"; }
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
runN() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
           bench.py --gpus $N --steps 100 --warmup 3 --skip-cpu --skip-variants "$@" > gpurun_out/r2_${tag}_n${N}_$name.json 2> gpurun_out/r2_${tag}_n${N}_$name.err; \
           line n${N}_$name < gpurun_out/r2_${tag}_n${N}_$name.json; grep -i "error\|Traceback" gpurun_out/r2_${tag}_n${N}_$name.err | head -3; \
           python -c "import json,sys; d=json.loads([l for l in open('gpurun_out/r2_${tag}_n${N}_$name.json') if l.startswith('{')][-1]); print(d['step_breakdown_ms']['step'], d['config']['l2_persist_bytes'])"; }
runN il16
runN il16_nopersist --no-l2-persist
runN il16_peer --exchange peer
runN il8 --block 8
