# A/B measurements on one B200 (scratch script for gpurun; results land in gpurun_out/)
set -x
timeout 200 python tools/bench_structural.py > gpurun_out/structural_fused4.json 2> gpurun_out/structural_fused4.err; cat gpurun_out/structural_fused4.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_compact --launch-skip 1 --launch-count 2 \
  -o gpurun_out/prof_compact -f python tools/bench_structural.py > gpurun_out/ncu_compact.log 2>&1
tail -2 gpurun_out/ncu_compact.log
