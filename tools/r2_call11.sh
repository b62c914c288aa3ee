set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "deviation_across_seeds" 2>&1 | grep -E "Error|assert|fired|gated|passed|failed" | head -20 | tee gpurun_out/r2_f_dev.log
timeout 600 python -m pytest tests/test_gpu_line32.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_f_tests32.log
run() { name=$1; shift; timeout 300 python bench.py --steps 30 --warmup 3 --skip-cpu --block 8 "$@" > gpurun_out/r2_f_$name.json 2> gpurun_out/r2_f_$name.err; python tools/bench_line.py f_$name < gpurun_out/r2_f_$name.json; }
run il8 --table-order interleaved
run dst8 --table-order dst
export ABNN_B200_LIB=$PWD/variants/lib_tune.so
ABNN_L2_ARRAYS=2 run il8_win40 --table-order interleaved
ABNN_L2_ARRAYS=1 run il8_win20 --table-order interleaved
ABNN_L2_MISS=1 run il8_missnormal --table-order interleaved
ABNN_L2_ARRAYS=3 ABNN_L2_RATIO=0.8 run il8_ratio08 --table-order interleaved
run il8_nopersist --table-order interleaved --no-l2-persist
run il8_novis --table-order interleaved --no-visits
