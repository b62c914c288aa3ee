#!/bin/bash
# iid sampler on the line32 kernel (sample_block 1): parity tests, then the kernel time at the 1B shape
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_line32.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_iid_tests.log
timeout 900 python -m pytest tests/test_gpu_equivalence.py tests/test_gpu_parity.py -m gpu -x -q -k "not full_size" 2>&1 | tail -8 | tee -a gpurun_out/r2_iid_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu --skip-variants --block 1 --table-order given > gpurun_out/r2_iid_bench.json 2> gpurun_out/r2_iid_bench.err
python tools/bench_line.py iid_line32 < gpurun_out/r2_iid_bench.json; tail -3 gpurun_out/r2_iid_bench.err
