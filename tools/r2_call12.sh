set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_line32.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_g_tests.log
run() { name=$1; shift; timeout 300 python bench.py --steps 30 --warmup 3 --skip-cpu --skip-variants "$@" > gpurun_out/r2_g_$name.json 2> gpurun_out/r2_g_$name.err; python tools/bench_line.py g_$name < gpurun_out/r2_g_$name.json; tail -2 gpurun_out/r2_g_$name.err; }
run il8 --block 8 --table-order interleaved
run il16 --block 16 --table-order interleaved
run dst8 --block 8 --table-order dst
run dst16 --block 16 --table-order dst
timeout 900 python -m pytest tests/test_gpu_equivalence.py -m gpu -x -q -s 2>&1 | tail -30 | tee gpurun_out/r2_g_equiv.log
