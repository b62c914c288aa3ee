"""profiles/r2_traffic.json from ncu captures (read on the CPU box): python tools/make_traffic.py <rep> <workload_key> [<rep> <key> ...]

Every entry carries the SHA-256 (first 16 hex digits) of the traversal kernel's source at the time the entry is made —
run this right after the capture, before touching traversal.cu / common.cuh: bench.py quotes `roofline.traffic` only when
the hash still matches (a stale capture yields traffic: null)."""
import csv, hashlib, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def source_hash():
    h = hashlib.sha256()
    for f in ("traversal.cu", "common.cuh"):
        h.update(open(os.path.join(ROOT, "abnn_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    out = {"captures": []}
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    args = sys.argv[1:]
    for rep, key in zip(args[0::2], args[1::2]):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, r = rows[0], rows[1], rows[2]
        get = lambda m: to_bytes(r[hdr.index(m)], units[hdr.index(m)])
        rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
        out["captures"].append({"kernel": r[hdr.index("Kernel Name")], "workload_key": key, "kernel_source_sha16": source_hash(),
                                "dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
                                "gpu_time_ms": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) *
                                               {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")], 1e-6),
                                "source": f"ncu --set full --clock-control none, one launch ({os.path.basename(rep)})"})
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
