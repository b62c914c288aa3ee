mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --steps 30 --warmup 3 --skip-cpu --skip-variants "$@" > gpurun_out/r2_j_$name.json 2> gpurun_out/r2_j_$name.err; python tools/bench_line.py j_$name < gpurun_out/r2_j_$name.json; tail -2 gpurun_out/r2_j_$name.err; }
timeout 900 python -m pytest tests/test_gpu_line32.py -m gpu -x -q 2>&1 | tail -3
ABNN_B200_LIB=$PWD/variants/lib_s7.so timeout 300 python -m pytest tests/test_gpu_line32.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do for blk in 8 16; do
  run base_il$blk.$rep --block $blk --table-order interleaved
  ABNN_B200_LIB=$PWD/variants/lib_s7.so run s7_il$blk.$rep --block $blk --table-order interleaved
done; done
