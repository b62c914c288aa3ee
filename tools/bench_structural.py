"""Structural step at scale (SURVEY.md §8d "compaction pass bytes"): stable prune-compaction + ordered growth + re-sort
on a 1B-synapse dst-sorted table. Prints one JSON line. Usage: python tools/bench_structural.py [--syn N]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from abnn_b200 import Brain, capi

ap = argparse.ArgumentParser()
ap.add_argument("--syn", type=int, default=1_000_000_000)
ap.add_argument("--hidden", type=int, default=5_000_000)
ap.add_argument("--events", type=int, default=150_000_000)
a = ap.parse_args()
p = capi.default_params(capi.PROFILE_B200)
p.n_hidden, p.n_syn, p.syn_capacity = a.hidden, a.syn, a.syn + (a.syn >> 6)
p.window_pre, p.refractory = 5 * a.events, 2 * a.events
p.w_prune, p.p_new, p.w_init = 0.05, 0.25, 0.1           # Beta(2,8): ~7 % of the weights are below 0.05
with Brain(p) as b:
    t0 = time.perf_counter(); b.init_graph(capi.GRAPH_ER_BETA, 1); b.sync(); t_init = time.perf_counter() - t0
    n = a.hidden + 512
    rng = np.random.default_rng(7)
    lf = np.zeros(n, np.uint64); idx = rng.choice(n, n // 4, replace=False)
    lf[idx] = rng.integers(a.events, 6 * a.events, len(idx)).astype(np.uint64)
    b.upload_timestamps(lf, None); b.clock = 6 * a.events; b.set_reward(0.01)
    st = b.run_pass(a.events)                              # fires stage growth candidates
    t0 = time.perf_counter(); ss = b.prune_and_grow(); t_struct = time.perf_counter() - t0
    t0 = time.perf_counter(); s2 = b.prune_and_grow(); t_prune_only = time.perf_counter() - t0   # nothing staged, nothing below w_prune: compaction sweep only
    st2 = b.run_pass(a.events)
    t0 = time.perf_counter(); ss3 = b.prune_and_grow(); t_struct2 = time.perf_counter() - t0     # steady state: spare table already allocated
    print(json.dumps({"n_syn": a.syn, "init_and_sort_s": t_init, "pass_fired": st.fired, "grown_staged": st.grown,
                      "pruned": ss.pruned, "appended": ss.appended, "dropped": ss.dropped, "n_after": ss.n_after,
                      "first_structural_step_ms": 1e3 * t_struct, "steady_structural_step_ms": 1e3 * t_struct2,
                      "steady_pruned": ss3.pruned, "steady_appended": ss3.appended, "compaction_sweep_ms": 1e3 * t_prune_only,
                      "compaction_sweep_GBps": 32.0 * ss.n_after / t_prune_only / 1e9}))
