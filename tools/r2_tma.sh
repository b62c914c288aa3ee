#!/bin/bash
# cp.async.bulk staging (tuning build variants/lib_tma.so, -DABNN_LINE_TMA=1) against the shipped LDGSTS staging, same box
mkdir -p gpurun_out
ABNN_B200_LIB=$PWD/variants/lib_tma.so timeout 600 python -m pytest tests/test_gpu_line32.py -m gpu -x -q 2>&1 | tail -4
run() { name=$1; shift; timeout 300 python bench.py --steps 30 --warmup 3 --skip-cpu --skip-variants "$@" > gpurun_out/r2_tma_$name.json 2> gpurun_out/r2_tma_$name.err; python tools/bench_line.py tma_$name < gpurun_out/r2_tma_$name.json; tail -1 gpurun_out/r2_tma_$name.err | cut -c1-200; }
for rep in 1 2; do
run ldgsts_il16.$rep
ABNN_B200_LIB=$PWD/variants/lib_tma.so run bulk_il16.$rep
run ldgsts_dst8.$rep --block 8 --table-order dst
ABNN_B200_LIB=$PWD/variants/lib_tma.so run bulk_dst8.$rep --block 8 --table-order dst
done
export ABNN_B200_LIB=$PWD/variants/lib_tma.so
bash tools/r2_ncu.sh tma_il16 k_traverse_line32 --skip-variants
