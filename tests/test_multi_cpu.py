"""N > 1 on the CPU (world_size 2, gloo): the host-side protocol of the dst-sharded path.

  * abnn_partition / abnn_event_share (host entry points of the C-ABI) against the oracle's.
  * NCCL unique-id hand-off through a torch.distributed group (abnn_b200.distributed).
  * Two processes, one oracle shard each, exchanging owned lastFired slices with a gloo all_gather and
    growth candidates with all_gather_object — the same steps the CUDA library performs with NCCL —
    must reproduce the single-process OracleWorld(2) bit for bit.
"""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from abnn_b200 import capi


def test_partition_and_event_share_match_oracle(oracle):
    lib, L = capi.load(), oracle.libb()
    rng = np.random.default_rng(0)
    for _ in range(300):
        n = int(rng.integers(1, 10_000_000)); w = int(rng.integers(1, 9)); r = int(rng.integers(0, w))
        a, b, c, d = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        assert lib.abnn_partition(n, w, r, C.byref(a), C.byref(b)) == 0
        assert L.ob_partition(n, w, r, C.byref(c), C.byref(d)) == 0
        assert (a.value, b.value) == (c.value, d.value)
        assert a.value <= b.value <= n and b.value - a.value <= -(-n // w)
    slices = [(lib.abnn_partition(1000, 3, r, C.byref(a), C.byref(b)), a.value, b.value) for r in range(3)]
    assert [s[1:] for s in slices] == [(0, 334), (334, 668), (668, 1000)]
    for _ in range(300):
        ev = int(rng.integers(0, 2**32)); cnt = rng.integers(0, 2**31, size=4)
        tot, acc = int(cnt.sum()), 0
        got = 0
        for k in range(4):
            f1, c1, f2, c2 = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
            lib.abnn_event_share(ev, tot, acc, int(cnt[k]), C.byref(f1), C.byref(c1))
            L.ob_event_share(ev, tot, acc, int(cnt[k]), C.byref(f2), C.byref(c2))
            assert (f1.value, c1.value) == (f2.value, c2.value) and f1.value == got
            got += c1.value; acc += int(cnt[k])
        assert got == (ev if tot else 0)          # shares tile the pass exactly


def test_philox_entry_point_matches_kat():
    from tests.test_oracle import PHILOX_KAT
    lib = capi.load()
    for ctr, key, want in PHILOX_KAT:
        o = (C.c_uint32 * 4)()
        lib.abnn_philox4x32((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), o)
        assert tuple(o) == want


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from abnn_b200 import distributed as D
    from oracle import pyoracle as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        uid = D.broadcast_unique_id()
        base = O.default_params(capi.PROFILE_NORTH_STAR, n_input=16, n_output=16, n_hidden=3001, n_syn=40_000,
                                exec_mode=capi.EXEC_SERIAL, window_pre=200_000, refractory=30_000, p_new=0.2,
                                w_prune=0.05, syn_capacity=30_000, sample_block=8)
        p = D.rank_params(base, rank, world)
        sh = O.OracleB(p)
        rng = np.random.default_rng(4)
        N = 16 + 16 + 3001
        syn = np.zeros(40_000, O.SYN_DTYPE)
        syn["src"] = rng.integers(0, N, 40_000); syn["dst"] = rng.integers(0, N, 40_000)
        syn["w"] = rng.uniform(0.02, 1.0, 40_000).astype(np.float32)
        pre = rng.integers(1, 50_000, N).astype(np.uint64)
        sh.upload_synapses(syn)                      # keeps the records whose dst this rank owns
        counts = [None] * world
        dist.all_gather_object(counts, sh.n_syn_local())
        sh.set_shard_counts(counts)
        sh.upload_timestamps(pre, None); sh.clock = 50_000; sh.set_reward(0.2)
        live, view = sh.live_view()
        lo, hi = rank * -(-N // world), min(N, (rank + 1) * -(-N // world))
        slice_len = -(-N // world)
        stats = []
        for it in range(4):
            vin = rng.random(16).astype(np.float32); exp = rng.random(16).astype(np.float32)
            sh.inject_inputs(vin, 1000.0); sh.teacher_force(exp, float(it & 1))
            st = sh.run_pass(30_000)
            # exchange: allgather of the owned lastFired slice (padded to equal slices), as NCCL does
            mine = np.zeros(slice_len, np.uint64); mine[:hi - lo] = live[lo:hi]
            outs = [torch.zeros(slice_len, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(outs, torch.from_numpy(mine.view(np.int64)))
            full = np.concatenate([o.numpy().view(np.uint64) for o in outs])[:N]
            view[:] = full
            stats.append((st.events, st.gated, st.fired, st.grown))
            if it == 1:                               # structural step: prune locally, exchange growth candidates
                sh.prune()
                cands = [None] * world
                dist.all_gather_object(cands, sh.grow_fetch())
                sh.grow_apply(np.concatenate(cands))
                dist.all_gather_object(counts, sh.n_syn_local())
                sh.set_shard_counts(counts)
        q.put((rank, uid, stats, sh.download_synapses().tobytes(), view.copy().tobytes(), sh.timestamps()[1][lo:hi].tobytes(), sh.clock))
    finally:
        dist.destroy_process_group()


def test_two_rank_protocol_equals_single_process_world(oracle):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in procs])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert res[0][1] == res[1][1] and len(res[0][1]) == 128 and any(res[0][1])       # same NCCL unique id on both ranks

    O = oracle
    base = O.default_params(capi.PROFILE_NORTH_STAR, n_input=16, n_output=16, n_hidden=3001, n_syn=40_000,
                            exec_mode=capi.EXEC_SERIAL, window_pre=200_000, refractory=30_000, p_new=0.2,
                            w_prune=0.05, syn_capacity=30_000, sample_block=8)
    world = O.OracleWorld(base, 2)
    rng = np.random.default_rng(4)
    N = 16 + 16 + 3001
    syn = np.zeros(40_000, O.SYN_DTYPE)
    syn["src"] = rng.integers(0, N, 40_000); syn["dst"] = rng.integers(0, N, 40_000)
    syn["w"] = rng.uniform(0.02, 1.0, 40_000).astype(np.float32)
    pre = rng.integers(1, 50_000, N).astype(np.uint64)
    world.upload_synapses(syn)
    for s in world.shards:
        s.upload_timestamps(pre, None); s.clock = 50_000; s.set_reward(0.2)
    per_rank = [[], []]
    for it in range(4):
        vin = rng.random(16).astype(np.float32); exp = rng.random(16).astype(np.float32)
        for s in world.shards:
            s.inject_inputs(vin, 1000.0); s.teacher_force(exp, float(it & 1))
        # run shard by shard to collect per-rank statistics, then exchange like ob_world_run_pass
        sts = [s.run_pass(30_000) for s in world.shards]
        for k, s in enumerate(world.shards):
            live, _ = s.live_view()
            for t in world.shards:
                _, view = t.live_view()
                lo, hi = k * -(-N // 2), min(N, (k + 1) * -(-N // 2))
                view[lo:hi] = live[lo:hi]
            per_rank[k].append((sts[k].events, sts[k].gated, sts[k].fired, sts[k].grown))
        if it == 1:
            st = world.prune_and_grow()
            assert st.pruned > 0 and st.appended > 0
    for k, s in enumerate(world.shards):
        rank, _, stats, syn_b, view_b, lv_b, clock = res[k]
        assert stats == per_rank[k]
        assert sum(x[1] for x in stats) > 100
        assert syn_b == s.download_synapses().tobytes()
        assert view_b == s.live_view()[1].tobytes()
        lo, hi = k * -(-N // 2), min(N, (k + 1) * -(-N // 2))
        assert lv_b == s.timestamps()[1][lo:hi].tobytes()
        assert clock == s.clock
