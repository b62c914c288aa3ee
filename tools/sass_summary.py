"""SASS evidence digest of the line kernels: python tools/sass_summary.py > profiles/r2_sass_line32.txt (after a build)."""
import collections, re, subprocess
OBJ = "abnn_b200/_obj/traversal.cu.o"
sass = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", OBJ], capture_output=True, text=True).stdout
filt = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
cur, body = None, collections.defaultdict(list)
for l in sass.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1); continue
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
    if cur and m:
        body[cur].append((m.group(1), m.group(2).strip()))
print("# SASS evidence for the data-movement choices of k_traverse_line32 (cuobjdump -sass of abnn_b200/_obj/traversal.cu.o, sm_100a,")
print("# nvcc 12.9 -O3 -fmad=false -lineinfo). Regenerate: python tools/sass_summary.py > profiles/r2_sass_line32.txt")
for name in sorted(body):
    if "line32" not in name:
        continue
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i[1]).split()[0] for i in body[name])
    pick = lambda pat: sum(v for k, v in ops.items() if re.match(pat, k))
    print(f"\n== {filt(name)}")
    print(f"   instructions {len(body[name])}; LDGSTS.E.BYPASS.128 (cp.async 16 B, L1 bypass) {ops.get('LDGSTS.E.BYPASS.128', 0)}; "
          f"ATOMG.*MAX (vis32 / fire32, result discarded = RED at L2) {pick(r'ATOMG\.E\.MAX')}; LDG.E (gate / fire words, ld.cg) {pick(r'LDG\.E')}; "
          f"STG.E (weight write-back) {pick(r'STG\.E')}; MATCH {pick(r'MATCH')}; SHFL {pick(r'SHFL')}; VOTE {pick(r'VOTE')}; "
          f"LDL/STL (local memory) {pick(r'LDL') + pick(r'STL')}; UTMALDG/UBLKCP (TMA) {pick(r'UTMA') + pick(r'UBLKCP')}")
for m in re.finditer(r"Function (\S*line32\S*):\s*\n\s*(REG:\d+[^\n]*)", res):
    print(f"   res-usage {filt(m.group(1))}: {m.group(2)}")
name = [n for n in body if "line32ILi1ELi0ELi16" in n][0]
print(f"\n-- excerpt of {filt(name)}: staging, visit RED, gate / fire loads, write-back, fire")
for a, t in [i for i in body[name] if re.search(r"LDGSTS|LDGDEPBAR|DEPBAR|ATOMG|STG\.E|LDG\.E", i[1])][:60]:
    print(f"   /*{a}*/ {t}")
