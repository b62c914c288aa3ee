mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --steps 30 --warmup 3 --skip-cpu --skip-variants "$@" > gpurun_out/r2_i_$name.json 2> gpurun_out/r2_i_$name.err; python tools/bench_line.py i_$name < gpurun_out/r2_i_$name.json; tail -2 gpurun_out/r2_i_$name.err; }
timeout 900 python -m pytest tests/test_gpu_line32.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for o in interleaved dst; do
  run base_$o --block 8 --table-order $o
  for lib in variants/lib_*.so; do v=$(basename $lib .so); v=${v#lib_}; ABNN_B200_LIB=$PWD/$lib run ${v}_$o --block 8 --table-order $o; done
done
ABNN_B200_LIB=$PWD/variants/lib_s7.so timeout 300 python -m pytest tests/test_gpu_line32.py -m gpu -x -q 2>&1 | tail -3
