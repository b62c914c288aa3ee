mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --steps 30 --warmup 3 --skip-cpu --skip-variants "$@" > gpurun_out/r2_h_$name.json 2> gpurun_out/r2_h_$name.err; python tools/bench_line.py h_$name < gpurun_out/r2_h_$name.json; tail -2 gpurun_out/r2_h_$name.err; }
timeout 600 python -m pytest tests/test_gpu_line32.py -m gpu -x -q 2>&1 | tail -3
for o in interleaved dst; do
  run base_$o --block 8 --table-order $o
  for lib in variants/lib_*.so; do v=$(basename $lib .so); v=${v#lib_}; ABNN_B200_LIB=$PWD/$lib run ${v}_$o --block 8 --table-order $o; done
done
run base_il16 --block 16 --table-order interleaved
timeout 900 python -m pytest tests/test_gpu_equivalence.py -m gpu -x -q -s 2>&1 | grep -E "passed|failed|line8|line16|Error" | tee gpurun_out/r2_h_equiv.log
