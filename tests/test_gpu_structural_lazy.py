"""Structural plasticity with periodic rebuilds (abnn_params.compact_every = K > 1, README.md:122-124 "remove, compact
periodically"): between two rebuilds a structural step marks the records that fell below w_prune dead IN PLACE (found
through the candidates the traversal kernels staged when they wrote such a weight — the step never sweeps the table) and
appends the grown synapses behind the table; every K-th step rebuilds the table (dead and pruned records leave, the tail
is merged into the table order). The oracle restates the same schedule with a full scan; every execution mode must match
it bit for bit in the deterministic modes: tables (including dead slots and the tail), stats, timestamps."""
import numpy as np
import pytest

from abnn_b200 import Brain, capi
from oracle import pyoracle as O
from tests.helpers import assert_same_state, assert_same_stats, random_graph

pytestmark = pytest.mark.gpu


def sstats(s):
    return (s.n_before, s.pruned, s.appended, s.dropped, s.n_after)


@pytest.mark.parametrize("mode,order", [(capi.EXEC_EXACT, capi.TABLE_DST_SORTED), (capi.EXEC_EXACT, capi.TABLE_DST_INTERLEAVED),
                                        (capi.EXEC_EXACT, capi.TABLE_AS_GIVEN), (capi.EXEC_SERIAL, capi.TABLE_DST_SORTED)])
def test_periodic_rebuild_bit_exact(mode, order):
    """A structural step after EVERY pass (BASELINE configs[4] regime), compact_every = 3, 8 steps: steps 0, 3, 6 rebuild,
    the others mark + append. Pruning and growth both active every step; dead records are sampled and skipped."""
    rng = np.random.default_rng(77)
    N, n = 20_000, 400_003
    syn = random_graph(rng, n, N, 0.15, 1.0, dst_lo=32)
    pre = rng.integers(1, 60_000, N).astype(np.uint64)
    events = 200_003 if mode == capi.EXEC_EXACT else 60_001
    p = O.default_params(capi.PROFILE_NORTH_STAR, n_input=32, n_output=32, n_hidden=N - 64, n_syn=n, exec_mode=mode, sample_block=8,
                         table_order=order, window_pre=400_000, refractory=20_000, p_new=0.2, w_prune=0.16, w_init=0.17,
                         syn_capacity=n + 200_000, compact_every=3)
    b, o = Brain(p), O.OracleB(p)
    for x in (b, o):
        x.upload_synapses(syn); x.upload_timestamps(pre, None); x.clock = 60_000; x.set_reward(0.1)
    dead_seen = tail_seen = 0
    for step in range(8):
        sb, so = b.run_pass(events), o.run_pass(events)
        assert_same_stats(sb, so, f"pass {step}")
        ssb, sso = b.prune_and_grow(), o.prune_and_grow()
        assert sstats(ssb) == sstats(sso), f"structural step {step}: {sstats(ssb)} vs {sstats(sso)}"
        tb, to = b.download_synapses(), o.download_synapses()
        assert tb.tobytes() == to.tobytes(), f"table after structural step {step}"
        dead = int((to["src"] == 0xFFFFFFFF).sum())
        if step % 3 == 0:
            assert dead == 0                                   # a rebuild leaves no dead record
            if order == capi.TABLE_DST_SORTED:
                assert np.all(np.diff(to["dst"].astype(np.int64)) >= 0)
        else:
            dead_seen += dead
            tail_seen += int(sso.appended)
            assert sso.n_after == sso.n_before + sso.appended   # slots only grow between rebuilds
    assert dead_seen > 100 and tail_seen > 100
    assert_same_state(b, o)


def test_line32_with_dead_records_and_tail_single_warp_exact():
    """The throughput kernel on a table with dead slots and an unsorted tail: 224-event passes (one warp orders its chunk
    exactly), a structural step every 5 passes, compact_every = 4 — bit-exact against the oracle, including the prune
    candidates the kernel stages when it writes a weight below w_prune."""
    rng = np.random.default_rng(5)
    N, n = 96, 96 * 300
    syn = random_graph(rng, n, N, 0.3, 1.0, dst_lo=16)
    p = O.default_params(capi.PROFILE_NORTH_STAR, exec_mode=capi.EXEC_PARALLEL, table_order=capi.TABLE_DST_INTERLEAVED,
                         n_input=8, n_output=8, n_hidden=N - 16, n_syn=n, sample_block=16, window_pre=10**9, refractory=300,
                         p_new=0.5, w_prune=0.3, w_init=0.31, syn_capacity=n + 8192, compact_every=4, a_ltd=0.05)
    b, o = Brain(p), O.OracleB(p)
    for x in (b, o):
        x.upload_synapses(syn); x.upload_timestamps(np.full(N, 1, np.uint64), None); x.clock = 1000; x.set_reward(0.2)
    pruned = grown = 0
    for q in range(60):
        assert_same_stats(b.run_pass(224), o.run_pass(224), f"pass {q}")
        if q % 5 == 4:
            ssb, sso = b.prune_and_grow(), o.prune_and_grow()
            assert sstats(ssb) == sstats(sso), f"structural step after pass {q}"
            assert b.download_synapses().tobytes() == o.download_synapses().tobytes()
            pruned += sso.pruned; grown += sso.appended
    assert pruned > 50 and grown > 50
    assert_same_state(b, o)


def test_state_file_resumes_inside_a_rebuild_cycle(tmp_path):
    """.bnn v2 written between two rebuilds (dead records, tail, staged prune candidates, step counter) resumes exactly."""
    rng = np.random.default_rng(9)
    N, n = 5000, 120_000
    syn = random_graph(rng, n, N, 0.15, 1.0, dst_lo=16)
    over = dict(n_input=16, n_output=16, n_hidden=N - 32, n_syn=n, exec_mode=capi.EXEC_EXACT, sample_block=8,
                table_order=capi.TABLE_DST_SORTED, window_pre=400_000, refractory=20_000, p_new=0.2, w_prune=0.16, w_init=0.17,
                syn_capacity=n + 60_000, compact_every=4)
    p = O.default_params(capi.PROFILE_NORTH_STAR, **over)
    b, o = Brain(p), O.OracleB(p)
    pre = rng.integers(1, 60_000, N).astype(np.uint64)
    for x in (b, o):
        x.upload_synapses(syn); x.upload_timestamps(pre, None); x.clock = 60_000; x.set_reward(0.1)
    path = str(tmp_path / "lazy.bnn2")
    for step in range(7):
        assert_same_stats(b.run_pass(80_001), o.run_pass(80_001), f"pass {step}")
        if step == 2:                                            # staged candidates of this pass travel in the file
            b.save_state(path)
            b.close()
            b = Brain(p)
            b.load_state(path)
        ssb, sso = b.prune_and_grow(), o.prune_and_grow()
        assert sstats(ssb) == sstats(sso), f"structural step {step}"
    assert_same_state(b, o)
    with Brain(O.default_params(capi.PROFILE_NORTH_STAR, **dict(over, compact_every=0))) as other:
        with pytest.raises(capi.AbnnError) as e:
            other.load_state(path)
        assert e.value.status == capi.ERR_SHAPE
