"""Generate tests/golden/*.json from the REFERENCE's own sources compiled verbatim (oracle/_ref).

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
  metal_kat.json       brain.metal (verbatim) on the SURVEY.md §8c KAT graph: per-pass budget, r-bar and
                       FNV-1a-64 of the synapse / lastFired arrays, naive and hold-clock sweeps.
  stimulus_filter.json FunctionalDataset frames (verbatim functional-dataset.cpp + the app's lambdas) and the
                       read-out chain (rate EMA -> verbatim RateFilter::process -> peak normalise) on a
                       seeded spike train, as float bit patterns.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402
from tests.test_oracle import kat_graph  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
assert O.have_ref(), "needs /root/reference to build oracle/_ref"

out = {}
for variant, hold in (("naive", False), ("hold_clock", True)):
    a = O.OracleA(kat_graph(), 64, reward=0.1, hold_clock=hold)
    passes = []
    for p in range(8):
        a.run_pass(1024)
        passes.append({"budget": int(a.st.budget), "rbar": float(np.float32(a.st.rbar)),
                       "syn": O.fnv1a64(a.syn.tobytes()), "lastF": O.fnv1a64(a.lastF.tobytes())})
    out[variant] = {"passes": passes, "lastF_head": a.lastF[:8].tolist(),
                    "w_spot": {str(i): float(a.syn["w"][i]) for i in (3, 100, 1023)}}
json.dump(out, open(os.path.join(HERE, "metal_kat.json"), "w"), indent=1)

R = O.RefPieces().L
frames = 12
h = R.refp_dataset_create(256, 256, 0.0009, 0.5)
ins, exps = [], []
v = np.zeros(256, np.float32)
for _ in range(frames):
    R.refp_dataset_next_input(h, v.ctypes.data); ins.append(v.view(np.uint32).tolist())
    R.refp_dataset_next_expected(h, v.ctypes.data); exps.append(v.view(np.uint32).tolist())
f = R.refp_filter_create(0.02, 1, 20)
rng = np.random.default_rng(5)
rate = np.zeros(256, np.float32); maxobs = np.float32(0.5); ro = []
for k in range(40):
    spikes = rng.random(256) < (0.2 + 0.6 * (k % 7) / 7)
    rate = (np.float32(0.5) * rate + np.float32(0.5) * spikes.astype(np.float32)).astype(np.float32)
    sm = np.zeros(256, np.float32)
    R.refp_filter_process(f, rate.ctypes.data, 256, 0.0009, sm.ctypes.data)
    maxobs = np.float32(max(maxobs, sm.max()) * np.float32(0.999))
    ro.append(np.minimum(sm / maxobs, np.float32(1.0)).astype(np.float32).view(np.uint32).tolist())
json.dump({"frames": frames, "input_bits": ins, "expected_bits": exps, "readout_bits": ro},
          open(os.path.join(HERE, "stimulus_filter.json"), "w"))
print("golden files written")
