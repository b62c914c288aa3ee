#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/exact_pass.py 2 > gpurun_out/r2_exact_plain.log 2>&1 && tail -1 gpurun_out/r2_exact_plain.log | cut -c1-300 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_exact_launches.csv python tools/exact_pass.py 1 > gpurun_out/r2_exact_ncu.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_exact_launches.csv') if l.startswith('"'))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
seq=[(r[ki].split('(')[0][-60:],int(r[vi])/1e6) for r in rows[1:]]
# print from the last k_exact_phase1 backwards to the previous end_pass
last=max(i for i,(k,v) in enumerate(seq) if 'k_exact_phase1' in k)
for k,v in seq[last-2:last+30]: print("%-62s %.3f ms"%(k,v))
P
