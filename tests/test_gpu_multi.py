"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests -m gpu -k multi`):
two ranks, one process and one GPU each, NCCL allgather of lastFired slices inside the library,
against the single-process OracleWorld(2) — bit-exact in SERIAL and EXACT execution, including the structural
step (local prune-compaction, allgathered growth candidates appended in event order; eager and compact_every > 1);
PARALLEL execution: exchanged gate words, lastVisited and statistics."""
import socket

import numpy as np
import pytest

from abnn_b200 import capi

pytestmark = pytest.mark.gpu

SCEN = dict(n_input=16, n_output=16, n_hidden=30_001, n_syn=400_000, window_pre=2_000_000, refractory=300_000,
            p_new=0.2, w_prune=0.05, syn_capacity=300_000, sample_block=8)


def _inputs():
    rng = np.random.default_rng(4)
    N = 16 + 16 + SCEN["n_hidden"]
    n = SCEN["n_syn"]
    from oracle.pyoracle import SYN_DTYPE
    syn = np.zeros(n, SYN_DTYPE)
    syn["src"] = rng.integers(0, N, n); syn["dst"] = rng.integers(0, N, n)
    syn["w"] = rng.uniform(0.02, 1.0, n).astype(np.float32)
    pre = rng.integers(1, 500_000, N).astype(np.uint64)
    frames = [(rng.random(16).astype(np.float32), rng.random(16).astype(np.float32)) for _ in range(4)]
    return N, syn, pre, frames


def _gpu_worker(rank, world, port, exec_mode, q, refractory=None, extra=None, struct_all=False):
    import torch
    import torch.distributed as dist
    from abnn_b200 import distributed as D
    from oracle import pyoracle as O
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        scen = dict(SCEN, refractory=refractory) if refractory else dict(SCEN)
        scen.update(extra or {})
        base = O.default_params(capi.PROFILE_NORTH_STAR, exec_mode=exec_mode, **scen)
        b = D.create_sharded_brain(base, device=rank)
        N, syn, pre, frames = _inputs()
        b.upload_synapses(syn)
        b.upload_timestamps(pre, None); b.clock = 500_000; b.set_reward(0.2)
        stats = []
        for it, (vin, exp) in enumerate(frames):
            b.inject_inputs(vin, 1000.0); b.teacher_force(exp, float(it & 1))
            st = b.run_pass(300_000)
            stats.append((st.events, st.gated, st.fired, st.grown))
            if it == 1 or struct_all:
                ss = b.prune_and_grow()
                stats.append((ss.pruned, ss.appended, ss.n_after, ss.dropped))
        info = b.info()
        lf, lv = b.timestamps()
        q.put((rank, stats, b.download_synapses().tobytes(), lf.tobytes(), lv[info.neuron_lo:info.neuron_hi].tobytes(),
               b.clock, info.n_syn_global, b.read_outputs().tobytes()))
        b.close()
    finally:
        dist.destroy_process_group()


def _gate_worker(rank, world, port, q, exchange=0, order=1, engine=False):
    """PARALLEL line kernel on a dst-sorted / interleaved table, sharded: after every pass the exchanged 32-bit gate
    words must equal clamp(window_pre - (clock - lastFired) + 1) of the replicated lastFired on every rank. exchange:
    abnn_params.exchange (NCCL allgather or peer-memory stores); engine: drive the passes through abnn_engine_step (on
    the peer exchange the sharded step is replayed from a CUDA graph from the third call on)."""
    import torch
    import torch.distributed as dist
    from abnn_b200 import distributed as D
    from oracle import pyoracle as O
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        scen = dict(SCEN, refractory=3_000, p_new=0.0, w_prune=0.0, table_order=order, exchange=exchange)
        base = O.default_params(capi.PROFILE_NORTH_STAR, exec_mode=capi.EXEC_PARALLEL, **scen)
        b = D.create_sharded_brain(base, device=rank)
        N, syn, pre, frames = _inputs()
        b.upload_synapses(syn)
        b.upload_timestamps(pre, None); b.clock = 500_000; b.set_reward(0.2)
        out = []
        info = b.info()
        for it, (vin, exp) in enumerate(frames + frames if engine else frames):
            if engine:
                b.engine_step(vin, exp, 1000.0, float(it & 1), 300_000)
                ev, fired = 150_000, 1000                 # no per-pass statistics on this path
            else:
                b.inject_inputs(vin, 1000.0); b.teacher_force(exp, float(it & 1))
                st = b.run_pass(300_000)
                ev, fired = st.events, st.fired
            words, valid = b.gate_words()
            lf, lv = b.timestamps()                       # collective: refreshes the 64-bit snapshot
            out.append((ev, fired, valid, words.tobytes(), lf.tobytes(), b.clock, b.read_outputs().tobytes(),
                        lv[info.neuron_lo:info.neuron_hi].tobytes()))
        q.put((rank, out))
        b.close()
    finally:
        dist.destroy_process_group()


def _oracle_visits(order, n_pass):
    """lastVisited of the two-shard oracle after n_pass passes of the gate-word scenario (an order-free max over the
    sampled events: the PARALLEL sharded run must reproduce it exactly, whatever the exchange)."""
    from oracle import pyoracle as O
    scen = dict(SCEN, refractory=3_000, p_new=0.0, w_prune=0.0, table_order=order)
    world = O.OracleWorld(O.default_params(capi.PROFILE_NORTH_STAR, exec_mode=capi.EXEC_SERIAL, **scen), 2)
    N, syn, pre, frames = _inputs()
    world.upload_synapses(syn)
    for s in world.shards:
        s.upload_timestamps(pre, None); s.clock = 500_000; s.set_reward(0.2)
    out = []
    for it, (vin, exp) in enumerate((frames + frames)[:n_pass]):
        for s in world.shards:
            s.inject_inputs(vin, 1000.0); s.teacher_force(exp, float(it & 1))
        world.run_pass(300_000)
        half = -(-N // 2)
        out.append([s.timestamps()[1][k * half:min(N, (k + 1) * half)].tobytes() for k, s in enumerate(world.shards)])
    return out


@pytest.mark.parametrize("exchange,order,engine", [(0, 1, False), (1, 1, False), (1, 2, False), (1, 2, True), (0, 2, True)],
                         ids=["nccl-sorted", "peer-sorted", "peer-interleaved", "peer-interleaved-graph", "nccl-interleaved-engine"])
def test_two_gpus_parallel_gate_word_exchange(exchange, order, engine):
    import multiprocessing as mp
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [ctx.Process(target=_gate_worker, args=(r, 2, port, q, exchange, order, engine)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=300) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    W = SCEN["window_pre"]
    fired = 0
    n_pass = 8 if engine else 4
    visits = _oracle_visits(order, n_pass)
    for it in range(n_pass):
        assert res[0][it][7] == visits[it][0] and res[1][it][7] == visits[it][1], f"lastVisited differs from the two-shard oracle after pass {it}"
        (ev0, f0, valid0, w0, lf0, c0, o0, _), (ev1, f1, valid1, w1, lf1, c1, o1, _) = res[0][it], res[1][it]
        assert valid0 and valid1, "the sharded PARALLEL exchange must deliver the gate words"
        assert w0 == w1 and lf0 == lf1 and c0 == c1 and o0 == o1, f"ranks disagree after pass {it}"
        lf = np.frombuffer(lf0, np.uint64).astype(np.int64)
        want = np.clip(W - (c0 - lf) + 1, 0, 0xFFFFFFFE).astype(np.uint32)
        got = np.frombuffer(w0, np.uint32)
        assert np.array_equal(got[32:], want[32:]), f"gate words differ from the lastFired snapshot after pass {it}"
        fired += f0 + f1
    assert fired > 1000


def _run_gpu(exec_mode, refractory=None, extra=None, struct_all=False):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, exec_mode, q, refractory, extra, struct_all)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=300) for _ in procs])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    return res


def _run_oracle(refractory=None, extra=None, struct_all=False):
    from oracle import pyoracle as O
    scen = dict(SCEN, refractory=refractory) if refractory else dict(SCEN)
    scen.update(extra or {})
    base = O.default_params(capi.PROFILE_NORTH_STAR, exec_mode=capi.EXEC_SERIAL, **scen)
    world = O.OracleWorld(base, 2)
    N, syn, pre, frames = _inputs()
    world.upload_synapses(syn)
    for s in world.shards:
        s.upload_timestamps(pre, None); s.clock = 500_000; s.set_reward(0.2)
    per_rank = [[], []]
    for it, (vin, exp) in enumerate(frames):
        for s in world.shards:
            s.inject_inputs(vin, 1000.0); s.teacher_force(exp, float(it & 1))
        sts = [s.run_pass(300_000) for s in world.shards]
        half = -(-N // 2)
        for k, s in enumerate(world.shards):
            live, _ = s.live_view()
            for t in world.shards:
                t.live_view()[1][k * half:min(N, (k + 1) * half)] = live[k * half:min(N, (k + 1) * half)]
            per_rank[k].append((sts[k].events, sts[k].gated, sts[k].fired, sts[k].grown))
        if it == 1 or struct_all:
            pruned = [s.prune() for s in world.shards]
            allc = np.concatenate([s.grow_fetch() for s in world.shards])
            for k, s in enumerate(world.shards):
                app, drop = s.grow_apply(allc)
                per_rank[k].append((pruned[k], app, s.n_syn_local(), drop))
            world._sync_counts()
    return world, per_rank, N


def test_two_gpus_serial_bit_exact():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    res = _run_gpu(capi.EXEC_SERIAL)
    world, per_rank, N = _run_oracle()
    half = -(-N // 2)
    for k, s in enumerate(world.shards):
        rank, stats, syn_b, lf_b, lv_b, clock, n_global, outs = res[k]
        assert stats == per_rank[k], (k, stats, per_rank[k])
        assert syn_b == s.download_synapses().tobytes()
        assert lf_b == s.live_view()[1].tobytes()                      # replicated view after the last exchange
        assert lv_b == s.timestamps()[1][k * half:min(N, (k + 1) * half)].tobytes()
        assert clock == s.clock
        assert n_global == sum(t.n_syn_local() for t in world.shards)
        assert outs == s.read_outputs().tobytes()
    assert sum(x[1] for x in per_rank[0][:2]) > 1000


def _assert_ranks_equal_oracle(res, world, per_rank, N):
    half = -(-N // 2)
    for k, s in enumerate(world.shards):
        rank, stats, syn_b, lf_b, lv_b, clock, n_global, outs = res[k]
        assert stats == per_rank[k], (k, stats, per_rank[k])
        assert syn_b == s.download_synapses().tobytes()
        assert lf_b == s.live_view()[1].tobytes()
        assert lv_b == s.timestamps()[1][k * half:min(N, (k + 1) * half)].tobytes()
        assert clock == s.clock
        assert n_global == sum(t.n_syn_local() for t in world.shards)
        assert outs == s.read_outputs().tobytes()


def test_two_gpus_exact_bit_exact():
    """EXACT execution on two dst-shards (per-destination buckets over the rank's own neuron slice) against the two-shard
    oracle: pass statistics, tables, both timestamp arrays, the structural step — bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    res = _run_gpu(capi.EXEC_EXACT)
    world, per_rank, N = _run_oracle()
    _assert_ranks_equal_oracle(res, world, per_rank, N)
    assert sum(x[1] for x in per_rank[0][:2]) > 1000


@pytest.mark.parametrize("mode", [capi.EXEC_SERIAL, capi.EXEC_EXACT], ids=["serial", "exact"])
def test_two_gpus_structural_step_every_pass_lazy_bit_exact(mode):
    """BASELINE configs[4] regime on two dst-shards: a structural step after EVERY pass with compact_every = 2 (steps 0 and 2
    rebuild, 1 and 3 mark dead in place and append behind the table; growth candidates allgathered, each rank appends the
    ones it owns in tick order) — per-step counts and final tables equal the two-shard oracle bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    extra = dict(compact_every=2, w_init=0.1, table_order=capi.TABLE_DST_SORTED)
    res = _run_gpu(mode, extra=extra, struct_all=True)
    world, per_rank, N = _run_oracle(extra=extra, struct_all=True)
    _assert_ranks_equal_oracle(res, world, per_rank, N)
    steps = [x for x in per_rank[0] if len(x) == 4][1::2]               # (pruned, appended, n_after, dropped) rows
    assert len(steps) == 4 and all(st[1] > 0 for st in steps) and any(st[0] > 0 for st in steps)


def test_two_gpus_parallel_statistical():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    # a 150k-event pass is smaller than what one B200 keeps in flight, so unordered execution sees no
    # intra-pass ordering at all; with a short refractory period order matters little and the counts agree
    res = _run_gpu(capi.EXEC_PARALLEL, refractory=3_000)
    world, per_rank, N = _run_oracle(refractory=3_000)
    for k in range(2):
        stats = res[k][1]
        for a, b in zip(stats, per_rank[k]):
            if len(a) == 4 and a[0] == b[0] and a[0] > 10_000:       # pass statistics rows (events equal)
                assert abs(a[1] - b[1]) <= 0.08 * b[1] + 50 and abs(a[2] - b[2]) <= 0.1 * b[2] + 50, (k, a, b)
