"""Freeze Oracle B's definition of the north-star semantics (the parts the reference leaves undefined: Philox edge
sampling, per-event uint64 clock, lastVisited, line sampler, dst-sorted table, pruning / synaptogenesis, dst sharding,
ER/Beta init — "parity unpinned" in DESIGN.md §3) as checksums: tests/golden/northstar_oracle.json.

    python tests/golden/make_northstar_golden.py

No reference implementation exists to pin these against, so the file does not make them "pinned"; it makes the
definition tamper-evident: tests/test_oracle.py::test_oracle_b_matches_frozen_northstar_checksums fails on the CPU if
oracle_b.cpp, its compile flags or the host toolchain ever change a result, and the GPU parity tests compare the CUDA
path with the same oracle. Cases (each: per-pass stats and SHA-256, first 16 hex digits, of table / lastFired / lastVisited):
  toy_reference   BASELINE configs[0]: 256/256/10k neurons, 1M synapses, reference graph (seed 1), iid Philox sampler,
                  sine input + teacher forcing every pass, 1M-event passes, growth on
  toy_line_sorted the throughput configuration: ER/Beta graph, line sampler (sample_block 8), dst-sorted table,
                  pruning + growth after every pass (BASELINE configs[4] regime)
  toy_two_shards  the same as toy_line_sorted on two dst-shards (OracleWorld): summed stats, per-shard checksums
  toy_b200_lazy   the layout bench.py runs (ABNN_PROFILE_B200: 16-record sample groups over the DST_INTERLEAVED table) with a
                  structural step after every pass and compact_every = 3: steps 0 and 3 rebuild, the others mark dead in
                  place and append behind the table (round 2)
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from abnn_b200 import capi  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TOY = dict(n_input=256, n_output=256, n_hidden=10_000, n_syn=1_000_000)
STAT = ("events", "gated", "fired", "candidates", "grown", "clock")


def sha(b):
    return hashlib.sha256(b).hexdigest()[:16]


def digest(o):
    lf, lv = o.timestamps()
    return {"syn": sha(o.download_synapses().tobytes()), "lastF": sha(lf.tobytes()), "lastV": sha(lv.tobytes())}


def stats(st):
    return {k: int(getattr(st, k)) for k in STAT}


def structural(ss):
    return {k: int(getattr(ss, k)) for k in ("n_before", "pruned", "appended", "dropped", "n_after")}


def warm(n, seed=9):
    rng = np.random.default_rng(seed)
    return rng.integers(1, 2_000_000, n).astype(np.uint64)


def case_toy_reference():
    p = O.default_params(capi.PROFILE_NORTH_STAR, **dict(TOY, exec_mode=capi.EXEC_SERIAL, window_pre=2_000_000, refractory=100_000,
                                                         seed=42, p_new=0.05, syn_capacity=1_020_000))
    o = O.OracleB(p)
    o.init_graph(capi.GRAPH_REFERENCE, 1)
    stim = O.Dataset(256, 256)
    out, even = [], False
    for i in range(3):
        vin, exp = stim.next_input(), stim.next_expected()
        o.inject_inputs(vin, 1000.0); o.teacher_force(exp, 1.0 if even else 0.0); even = not even
        if i == 1:
            o.set_reward(0.05)
        st = o.run_pass(1_000_000)
        rates = o.readout_filtered(exp)
        out.append({"stats": stats(st), "readout": sha(rates.tobytes()), **digest(o)})
    out.append({"structural": structural(o.prune_and_grow()), **digest(o)})
    return out


def line_params(**kw):
    return O.default_params(capi.PROFILE_NORTH_STAR, **dict(TOY, exec_mode=capi.EXEC_SERIAL, sample_block=8, table_order=capi.TABLE_DST_SORTED,
                                                            window_pre=2_000_000, refractory=100_000, seed=42, p_new=0.1, w_prune=0.03,
                                                            w_init=0.1, syn_capacity=1_100_000, **kw))


def case_toy_line_sorted():
    o = O.OracleB(line_params())
    o.init_graph(capi.GRAPH_ER_BETA, 7)
    o.upload_timestamps(warm(10_512), None); o.clock = 2_000_000; o.set_reward(0.02)
    out = []
    for i in range(3):
        st = o.run_pass(1_000_000)
        ss = o.prune_and_grow()
        out.append({"stats": stats(st), "structural": structural(ss), **digest(o)})
    return out


def case_toy_two_shards():
    w = O.OracleWorld(line_params(), 2)
    w.init_graph(capi.GRAPH_ER_BETA, 7)
    for s in w.shards:
        s.upload_timestamps(warm(10_512), None); s.clock = 2_000_000; s.set_reward(0.02)
    out = []
    for i in range(3):
        st = w.run_pass(1_000_000)
        ss = w.prune_and_grow()
        out.append({"stats": stats(st), "structural": structural(ss), "shards": [digest(s) for s in w.shards]})
    return out


def case_toy_b200_lazy():
    p = O.default_params(capi.PROFILE_B200, **dict(TOY, exec_mode=capi.EXEC_SERIAL, window_pre=2_000_000, refractory=100_000, seed=42,
                                                   p_new=0.1, w_prune=0.03, w_init=0.1, syn_capacity=1_100_000, compact_every=3))
    assert p.sample_block == 16 and p.table_order == capi.TABLE_DST_INTERLEAVED
    o = O.OracleB(p)
    o.init_graph(capi.GRAPH_ER_BETA, 7)
    o.upload_timestamps(warm(10_512), None); o.clock = 2_000_000; o.set_reward(0.02)
    out = []
    for i in range(5):
        st = o.run_pass(1_000_000)
        ss = o.prune_and_grow()
        out.append({"stats": stats(st), "structural": structural(ss), **digest(o)})
    return out


CASES = {"toy_reference": case_toy_reference, "toy_line_sorted": case_toy_line_sorted, "toy_two_shards": case_toy_two_shards,
         "toy_b200_lazy": case_toy_b200_lazy}

if __name__ == "__main__":
    out = {name: fn() for name, fn in CASES.items()}
    json.dump(out, open(os.path.join(HERE, "northstar_oracle.json"), "w"), indent=1)
    for name, rows in out.items():
        print(name, json.dumps(rows[-1])[:200])
