#!/bin/bash
# EXACT execution: parity tests that run it, then the pass time and launch list at the bench shape
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_equivalence.py tests/test_gpu_structural_lazy.py tests/test_host_cpp.py -m gpu -q -x 2>&1 | tail -6
bash tools/r2_exact.sh
