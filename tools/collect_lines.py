"""Collect bench lines from gpurun_out/ into profiles/r2_bench_lines.jsonl: python tools/collect_lines.py <label>=<file> ..."""
import json, sys
out = []
for arg in [a for a in sys.argv[1:] if a != "--append"]:
    label, path = arg.split("=", 1)
    try:
        line = [l for l in open(path) if l.startswith("{")][-1]
    except Exception as e:
        print("skip", arg, e); continue
    d = json.loads(line)
    d = {"label": label, **d}
    out.append(json.dumps(d))
open("profiles/r2_bench_lines.jsonl", "a" if "--append" in sys.argv else "w").write("\n".join(out) + "\n")
print(len(out), "lines")
