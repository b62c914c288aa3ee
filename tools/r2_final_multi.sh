#!/bin/bash
# final multi-GPU call: bash tools/r2_final_multi.sh <N on the box> ; runs the 2-GPU tests (N = 2) and the bench at every N' <= N asked for
NBOX=$1; shift
mkdir -p gpurun_out
runN() { N=$1; name=$2; shift 2; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
           bench.py --gpus $N --warmup 3 --skip-cpu "$@" > gpurun_out/r2_final_n${N}_$name.json 2> gpurun_out/r2_final_n${N}_$name.err; \
           python tools/bench_line.py n${N}_$name < gpurun_out/r2_final_n${N}_$name.json; grep -i "error\|Traceback" gpurun_out/r2_final_n${N}_$name.err | head -3; \
           python -c "import json,sys; d=json.loads([l for l in open('gpurun_out/r2_final_n${N}_$name.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['step_breakdown_ms']['traverse'], (d.get('parity') or {}).get('ok'), d.get('structural'))"; }
if [ "$NBOX" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_final_multi_tests.log
  runN 2 nccl --steps 100
  runN 2 peer --steps 100 --exchange peer --skip-variants
else
  runN 8 nccl --steps 100
  runN 8 peer --steps 100 --exchange peer --skip-variants
  NCCL_ALGO=NVLS runN 8 nccl_algo_nvls --steps 100 --skip-variants          # deployment knobs of NCCL itself, not of the library
  NCCL_PROTO=LL128 runN 8 nccl_proto_ll128 --steps 100 --skip-variants
  runN 4 nccl --steps 100
  runN 8 structural --structural --steps 128
fi
