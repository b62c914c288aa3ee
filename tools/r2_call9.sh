set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_line32.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_d_tests.log
timeout 900 python -m pytest tests/test_gpu_equivalence.py -m gpu -x -q -s 2>&1 | tail -40 | tee gpurun_out/r2_d_equiv.log
for o in interleaved dst; do for blk in 8 16; do
  timeout 300 python bench.py --steps 30 --warmup 3 --skip-cpu --block $blk --table-order $o > gpurun_out/r2_d_${o}_b$blk.json 2> gpurun_out/r2_d_${o}_b$blk.err; python tools/bench_line.py d_${o}_b$blk < gpurun_out/r2_d_${o}_b$blk.json
done; done
