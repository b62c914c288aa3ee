#!/bin/bash
# multi-GPU call: bash tools/r2_multi.sh <N> <tag>   (gpurun --gpus N)
N=$1; tag=$2
mkdir -p gpurun_out
line() { python tools/bench_line.py "$1"; }
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_${tag}_multi_tests.log
fi
runN() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
           bench.py --gpus $N --steps 100 --warmup 3 --skip-cpu "$@" > gpurun_out/r2_${tag}_n${N}_$name.json 2> gpurun_out/r2_${tag}_n${N}_$name.err; \
           line n${N}_$name < gpurun_out/r2_${tag}_n${N}_$name.json; tail -3 gpurun_out/r2_${tag}_n${N}_$name.err | cut -c1-300; }
runN nccl --skip-variants
runN peer --exchange peer
runN nccl_parity
