# A/B measurements on one B200 (scratch script for gpurun; results land in gpurun_out/)
set -x
run() {  # name, extra bench args
  name=$1; shift
  timeout 200 python bench.py --steps 20 --warmup 3 --skip-cpu "$@" > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python tools/bench_line.py $name < gpurun_out/ab_$name.json
}
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ABNN_B200_LIB=variants/lib_v9.so run v9b
run vis32
ABNN_L2_ARRAYS=3 run vis32_w80
