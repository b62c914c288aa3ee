mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --warmup 2 --steps 6 --skip-cpu --skip-variants --block 1 --table-order given "$@" > gpurun_out/r2_s_$name.json 2> gpurun_out/r2_s_$name.err; python tools/bench_line.py s_$name < gpurun_out/r2_s_$name.json; tail -1 gpurun_out/r2_s_$name.err; }
for lib in t3 t4 t5 t6; do for u in 4 8; do
  ABNN_B200_LIB=$PWD/variants/lib_$lib.so ABNN_TRAV_U=$u run ${lib}_u$u
done; done
ABNN_B200_LIB=$PWD/variants/lib_t4.so ABNN_TRAV_U=2 run t4_u2
