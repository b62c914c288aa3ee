#!/bin/bash
# one GPU call of the round: tests, then benches; results in gpurun_out/r2_*
set -x
mkdir -p gpurun_out
tag=$1; shift
line() { python tools/bench_line.py "$1"; }
timeout 600 python -m pytest tests/test_gpu_line32.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_${tag}_tests32.log
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_line32.py 2>&1 | tail -8 | tee gpurun_out/r2_${tag}_tests.log
for blk in 8 16; do
  timeout 300 python bench.py --steps 50 --warmup 3 --skip-cpu --block $blk > gpurun_out/r2_${tag}_b$blk.json 2> gpurun_out/r2_${tag}_b$blk.err; line ${tag}_b$blk < gpurun_out/r2_${tag}_b$blk.json
done
