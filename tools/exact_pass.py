"""A few EXACT-mode passes at the bench shape (for `ncu --metrics gpu__time_duration.sum`): python tools/exact_pass.py [passes] [bench args]"""
import sys, json
sys.path.insert(0, ".")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sys.argv = sys.argv[:1] + ["--skip-cpu", "--skip-variants"] + sys.argv[2:]
import bench
from abnn_b200 import capi
args = bench.parse()
print(json.dumps(bench.sub_record(args, "exact", 6541.1, passes=n, exec_mode=capi.EXEC_EXACT)))
