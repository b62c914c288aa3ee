#!/bin/bash
# ncu captures of the traversal kernel: bash tools/r2_ncu.sh <tag> <kernel regex> <bench args...>
set -x
mkdir -p gpurun_out
tag=$1; shift; pat=$1; shift
CMD="python bench.py --steps 3 --warmup 3 --skip-cpu $*"
$CMD > gpurun_out/r2_${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$pat -s 6 -c 1 -o gpurun_out/r2_${tag}_prof -f $CMD > gpurun_out/r2_${tag}_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_${tag}_launches.csv $CMD > gpurun_out/r2_${tag}_ncu2.log 2>&1
tail -3 gpurun_out/r2_${tag}_ncu.log
