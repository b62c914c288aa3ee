#!/bin/bash
# iid sampler (sample_block 1): chunk-size variants of k_traverse_line32 and the register-staged kernel, same box
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --steps 20 --warmup 3 --skip-cpu --skip-variants --block 1 --table-order given > gpurun_out/r2_iidv_$name.json 2> gpurun_out/r2_iidv_$name.err; python tools/bench_line.py iid_$name < gpurun_out/r2_iidv_$name.json; }
run s8
ABNN_B200_LIB=$PWD/variants/lib_s6.so run s6
ABNN_B200_LIB=$PWD/variants/lib_s4.so run s4
ABNN_IID_LEGACY=1 ABNN_B200_LIB=$PWD/variants/lib_s6.so run legacy
