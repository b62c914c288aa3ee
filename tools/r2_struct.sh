#!/bin/bash
# structural plasticity every pass (configs[4] per-GPU load on one GPU): lazy-mode tests, then compact_every 16 / 64 / 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_structural_lazy.py -m gpu -x -q 2>&1 | tail -5
run() { name=$1; shift; timeout 300 python bench.py --warmup 3 --skip-cpu --skip-variants "$@" > gpurun_out/r2_s_$name.json 2> gpurun_out/r2_s_$name.err; python tools/bench_line.py s_$name < gpurun_out/r2_s_$name.json; tail -2 gpurun_out/r2_s_$name.err; python -c "import json; d=json.loads([l for l in open('gpurun_out/r2_s_$name.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d.get('structural'))"; }
run struct_k16 --structural --steps 48
run struct_k64 --structural --steps 128 --compact-every 64
run struct_k1 --structural --steps 10 --compact-every 1
