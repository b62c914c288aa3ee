import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline", {})
    print("%-28s value %.2fe9 ev/s  g=%.3f  kernel %.3f ms  frac %.3f sector %.3f  e2e %.2fe9  clocks %s" % (
        sys.argv[1] if len(sys.argv) > 1 else "", d["value"] / 1e9, d.get("gated_fraction", -1), r.get("kernel_ms", -1),
        r.get("frac", -1), r.get("sector_level_frac", -1), d["e2e"]["value"] / 1e9, (d.get("clocks") or {}).get("sm_mhz")))
