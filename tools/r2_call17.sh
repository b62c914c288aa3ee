mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --steps 50 --warmup 3 --skip-cpu --skip-variants "$@" > gpurun_out/r2_l_$name.json 2> gpurun_out/r2_l_$name.err; python tools/bench_line.py l_$name < gpurun_out/r2_l_$name.json; tail -2 gpurun_out/r2_l_$name.err; }
timeout 900 python -m pytest tests/test_gpu_line32.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
run full
run half --syn 500000000 --hidden 2500000 --events 75000000
run eighth --syn 125000000 --hidden 625000 --events 18750000
run half_syn_only --syn 500000000
run half_events_only --events 75000000
