#!/bin/bash
# ncu captures of the shipped kernels (single GPU): headline (block 16, interleaved), round-1 layout (block 8, dst-sorted), iid
bash tools/r2_ncu.sh fin_il16 k_traverse_line32 --skip-variants
bash tools/r2_ncu.sh fin_dst8 k_traverse_line32 --skip-variants --block 8 --table-order dst
CMD="python bench.py --steps 2 --warmup 2 --skip-cpu --skip-variants --block 1 --table-order given"
$CMD > gpurun_out/r2_fin_iid_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_traverse_parallel -s 4 -c 1 -o gpurun_out/r2_fin_iid_prof -f $CMD > gpurun_out/r2_fin_iid_ncu.log 2>&1
tail -2 gpurun_out/r2_fin_iid_ncu.log
