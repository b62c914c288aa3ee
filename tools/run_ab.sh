# Verification + measurements on one B200 (scratch script for gpurun; results land in gpurun_out/)
set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 200 python tools/bench_structural.py > gpurun_out/structural_final.json 2> gpurun_out/structural_final.err; cat gpurun_out/structural_final.json
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; python tools/bench_line.py final < gpurun_out/bench_final.json; tail -c 600 gpurun_out/bench_final.json
