"""Collect bench lines from gpurun_out/ into profiles/r2_bench_lines.jsonl: python tools/collect_lines.py <label>=<file> ..."""
import json, sys
out = []
for arg in sys.argv[1:]:
    label, path = arg.split("=", 1)
    try:
        line = [l for l in open(path) if l.startswith("{")][-1]
    except Exception as e:
        print("skip", arg, e); continue
    d = json.loads(line)
    d = {"label": label, **d}
    out.append(json.dumps(d))
open("profiles/r2_bench_lines.jsonl", "w").write("\n".join(out) + "\n")
print(len(out), "lines")
