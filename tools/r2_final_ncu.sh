#!/bin/bash
# ncu captures of the shipped kernels (single GPU): headline (block 16, interleaved), round-1 layout (block 8, dst-sorted), iid
bash tools/r2_ncu.sh fin_il16 k_traverse_line32 --skip-variants
bash tools/r2_ncu.sh fin_dst8 k_traverse_line32 --skip-variants --block 8 --table-order dst
bash tools/r2_ncu.sh fin_iid k_traverse_line32 --skip-variants --block 1 --table-order given
