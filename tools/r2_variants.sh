#!/bin/bash
# bench the default library and every variants/lib_*.so (tuning builds): bash tools/r2_variants.sh <tag> <bench args...>
mkdir -p gpurun_out
tag=$1; shift
line() { python tools/bench_line.py "$1"; }
run() { name=$1; shift; timeout 300 python bench.py --steps 30 --warmup 3 --skip-cpu "$@" > gpurun_out/r2_${tag}_$name.json 2> gpurun_out/r2_${tag}_$name.err; line ${tag}_$name < gpurun_out/r2_${tag}_$name.json; }
for blk in 8 16; do
  run base_b$blk --block $blk "$@"
  for lib in variants/lib_*.so; do
    [ -f "$lib" ] || continue
    v=$(basename $lib .so); v=${v#lib_}
    ABNN_B200_LIB=$PWD/$lib run ${v}_b$blk --block $blk "$@"
  done
done
