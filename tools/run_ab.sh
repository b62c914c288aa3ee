# Verification + measurements on one B200 (scratch script for gpurun; results land in gpurun_out/)
set -x
timeout 400 python -m pytest tests -m gpu -x -q -k "structural or prune or dst_sorted or bnn_v2 or engine_loop or toy" 2>&1 | tail -3
timeout 200 python tools/bench_structural.py > gpurun_out/structural_radix2.json 2> gpurun_out/structural_radix2.err; cat gpurun_out/structural_radix2.json
