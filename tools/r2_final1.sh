#!/bin/bash
# final single-GPU call of the round: tests, smoke, default bench, reference arm
mkdir -p gpurun_out
free -g | head -2; nproc
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2_final_smoke.log
timeout 600 python bench.py > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err; python tools/bench_line.py final_n1 < gpurun_out/r2_final_n1.json; tail -2 gpurun_out/r2_final_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; cut -c1-600 gpurun_out/r2_final_ref.json; tail -2 gpurun_out/r2_final_ref.err
