mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_equivalence.py --deselect tests/test_gpu_multi.py 2>&1 | tail -6
run() { name=$1; shift; timeout 300 python bench.py --warmup 3 --skip-cpu --skip-variants "$@" > gpurun_out/r2_r_$name.json 2> gpurun_out/r2_r_$name.err; python tools/bench_line.py r_$name < gpurun_out/r2_r_$name.json; tail -2 gpurun_out/r2_r_$name.err; }
run iid --steps 10 --block 1 --table-order given
run b4 --steps 10 --block 4 --table-order interleaved
run struct_k16 --structural --steps 48
python -c "import json; d=json.loads([l for l in open('gpurun_out/r2_r_struct_k16.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d.get('structural'))"
