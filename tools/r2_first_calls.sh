#!/bin/bash
# Next round's first GPU calls (everything here was written at the end of round 1 without GPU time left to run it).
#   gpurun --timeout 900 -- 'bash tools/r2_first_calls.sh one'          # one B200, ~3 min
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/r2_first_calls.sh two'  # two B200s, ~2 min (charged twice)
# Results land in gpurun_out/r2_*.
set -x
mkdir -p gpurun_out
line() { python tools/bench_line.py "$1"; }
case "$1" in
one)
  timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2_tests.log
  # 1. where the ceiling is: the line sampler's memory streams without its arithmetic
  (cd tools && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o probe_line probe_line.cu) && timeout 120 tools/probe_line | tee gpurun_out/r2_probe_line.txt
  # 2. iid sampler (sample_block 1) with and without the 32-bit gate words
  timeout 200 python bench.py --steps 20 --warmup 3 --skip-cpu --block 1 > gpurun_out/r2_iid.json 2> gpurun_out/r2_iid.err; line iid < gpurun_out/r2_iid.json
  ABNN_IID_SLACK=1 timeout 200 python bench.py --steps 20 --warmup 3 --skip-cpu --block 1 > gpurun_out/r2_iid_slack.json 2> gpurun_out/r2_iid_slack.err; line iid_slack < gpurun_out/r2_iid_slack.json
  ABNN_IID_SLACK=1 timeout 300 python -m pytest tests -m gpu -x -q -k "parallel or block" 2>&1 | tail -3 | tee gpurun_out/r2_tests_iid_slack.log
  ;;
two)
  # 3. peer-memory exchange instead of NCCL: parity first, then the N=2 bench with and without, then with the captured step
  ABNN_P2P_EXCHANGE=1 timeout 250 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2_tests_p2p.log
  run2() { name=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
             bench.py --gpus 2 --steps 100 --warmup 3 --skip-cpu > gpurun_out/r2_n2_$name.json 2> gpurun_out/r2_n2_$name.err; line n2_$name < gpurun_out/r2_n2_$name.json; }
  run2 nccl
  ABNN_P2P_EXCHANGE=1 run2 p2p
  ABNN_P2P_EXCHANGE=1 ABNN_P2P_GRAPH=1 run2 p2p_graph
  ;;
*) echo "usage: $0 one|two"; exit 2;;
esac
