"""GPU parity tests of k_traverse_line32 (abnn_b200/csrc/traversal.cu): the line sampler on 32-bit pass-relative
timestamps — the kernel bench.py times. It runs when the gate words are in use, the clock ticks per event, there is no
spike budget and the refractory period covers at least one chunk (256 ticks x world size); every test here is inside
those preconditions (the 64-bit line kernel k_traverse_line, which serves everything else, is covered by
tests/test_gpu_parity.py). All through the C-ABI, against Oracle B on the same seeded inputs.
"""
import numpy as np
import pytest

from abnn_b200 import Brain, capi
from oracle import pyoracle as O
from tests.helpers import assert_same_state, assert_same_stats, random_graph

pytestmark = pytest.mark.gpu


def pair(**over):
    p = O.default_params(capi.PROFILE_NORTH_STAR, exec_mode=capi.EXEC_PARALLEL, table_order=capi.TABLE_DST_SORTED, **over)
    return Brain(p), O.OracleB(p)


@pytest.mark.parametrize("block", [1, 8, 16])
def test_single_warp_chains_bit_exact(block):
    """One chunk (224 events <= one chunk of the kernel = one warp) per pass over a dst-sorted table with 64 destinations x 4096 synapses: every
    same-destination dependency of a pass is inside the warp, so PARALLEL must equal the serial oracle bit for bit —
    fire decisions, both timestamp arrays, weights, staged growth — over 60 passes. The refractory period (300 ticks)
    spans passes, so the 32-bit fire words are rebuilt from / folded into the 64-bit array with live content each pass."""
    rng = np.random.default_rng(33 + block)
    N, n = 64, 64 * 4096
    syn = np.zeros(n, O.SYN_DTYPE)
    syn["src"] = rng.integers(0, N, n); syn["dst"] = np.repeat(np.arange(N), 4096)
    syn["w"] = rng.uniform(0.3, 1.0, n).astype(np.float32)
    b, o = pair(n_input=8, n_output=8, n_hidden=N - 16, n_syn=n, sample_block=block, window_pre=10**9, refractory=300,
                p_new=0.3, syn_capacity=n + 4096)
    for x in (b, o):
        x.upload_synapses(syn); x.upload_timestamps(np.full(N, 1, np.uint64), None); x.clock = 1000; x.set_reward(0.2)
    fired = 0
    for p in range(60):
        sb, so = b.run_pass(224), o.run_pass(224)
        assert_same_stats(sb, so, f"pass {p}")
        fired += so.fired
    assert fired > 300
    assert_same_state(b, o)
    sb, so = b.prune_and_grow(), o.prune_and_grow()
    assert (sb.appended, sb.n_after) == (so.appended, so.n_after) and so.appended > 20
    assert_same_state(b, o, "after growth")


@pytest.mark.parametrize("block", [1, 8, 16])
def test_duplicate_lines_in_a_chunk_bit_exact(block):
    """A 64-line table (512 records, 16 destinations x 4 lines): every 256-event chunk draws 32 lines out of 64, so
    most chunks hold the same line twice or more (224-event passes: one chunk). The later copy must see the weights and the fires of the earlier one
    (the kernel cuts a dense step in front of a repeated line and re-reads the weights): bit-exact over 100 passes.
    block = 1 is the iid sampler on the same kernel (one 16-byte record per draw): 224 draws out of 512 records."""
    rng = np.random.default_rng(5 + block)
    N, n = 32, 512
    syn = np.zeros(n, O.SYN_DTYPE)
    syn["src"] = rng.integers(0, N, n); syn["dst"] = 16 + np.repeat(np.arange(16), 32)
    syn["w"] = rng.uniform(0.2, 1.0, n).astype(np.float32)
    b, o = pair(n_input=8, n_output=8, n_hidden=N - 16, n_syn=n, sample_block=block, window_pre=10**9, refractory=700)
    for x in (b, o):
        x.upload_synapses(syn); x.upload_timestamps(np.full(N, 1, np.uint64), None); x.clock = 5000; x.set_reward(-0.3)
    fired = gated = 0
    for p in range(100):
        sb, so = b.run_pass(224), o.run_pass(224)
        assert_same_stats(sb, so, f"pass {p}")
        fired += so.fired; gated += so.gated
    assert fired > 100 and gated > 500
    assert_same_state(b, o)


def test_iid_sampler_large_table_single_warp_bit_exact():
    """sample_block = 1 over a table of 2^24 records — the size from which the kernel no longer looks for records drawn
    twice inside a chunk (probability 1.5e-3 per 224-event pass here, 3e-5 per chunk at 1B records; the seed is chosen so
    that no pass does — such a pair is two concurrent events on one weight, the race PARALLEL execution has anyway):
    one chunk per pass, bit-exact against the oracle over 40 passes, table as generated (no order)."""
    rng = np.random.default_rng(1234)
    N, n = 256, 1 << 24
    syn = random_graph(rng, n, N, 0.3, 1.0, dst_lo=16)
    def draws_a_record_twice(seed):
        for q in range(40):
            edges = set()
            for e in range(q * 224, q * 224 + 224):
                r = O.philox([e & 0xFFFFFFFF, e >> 32, 0, 0], [seed & 0xFFFFFFFF, seed >> 32])
                edges.add((((r[0] << 32) | r[1]) * n) >> 64)
            if len(edges) < 224:
                return True
        return False
    seed = next(s for s in range(42, 60) if not draws_a_record_twice(s))      # 42 repeats record 6483210 in pass 2
    p = O.default_params(capi.PROFILE_NORTH_STAR, exec_mode=capi.EXEC_PARALLEL, table_order=capi.TABLE_AS_GIVEN, seed=seed,
                         n_input=16, n_output=16, n_hidden=N - 32, n_syn=n, sample_block=1, window_pre=10**9, refractory=300)
    b, o = Brain(p), O.OracleB(p)
    for x in (b, o):
        x.upload_synapses(syn); x.upload_timestamps(np.full(N, 1, np.uint64), None); x.clock = 1000; x.set_reward(0.1)
    fired = gated = 0
    for q in range(40):
        sb, so = b.run_pass(224), o.run_pass(224)
        assert_same_stats(sb, so, f"pass {q}")
        fired += so.fired; gated += so.gated
    assert fired > 200 and gated > 1000
    assert_same_state(b, o)


@pytest.mark.parametrize("block", [8, 16])
def test_interleaved_table_single_warp_bit_exact(block):
    """DST_INTERLEAVED table (8 adjacent destinations per line): a 256-event pass is one warp, which orders its events
    exactly; random in-degrees (ragged groups), growth on. Bit-exact against the oracle over 60 passes, then a
    structural step (the interleaved order is re-derived on both sides)."""
    rng = np.random.default_rng(71 + block)
    N, n = 96, 96 * 700
    syn = random_graph(rng, n, N, 0.3, 1.0, dst_lo=13)
    p = O.default_params(capi.PROFILE_NORTH_STAR, exec_mode=capi.EXEC_PARALLEL, table_order=capi.TABLE_DST_INTERLEAVED,
                         n_input=8, n_output=8, n_hidden=N - 16, n_syn=n, sample_block=block, window_pre=10**9, refractory=300,
                         p_new=0.3, w_prune=0.31, syn_capacity=n + 4096)
    b, o = Brain(p), O.OracleB(p)
    for x in (b, o):
        x.upload_synapses(syn); x.upload_timestamps(np.full(N, 1, np.uint64), None); x.clock = 1000; x.set_reward(0.2)
    assert b.download_synapses().tobytes() == o.download_synapses().tobytes(), "interleaved order differs after upload"
    t = b.download_synapses()["dst"]
    assert len(set(t[8000:8008].tolist())) == 8                         # a line in the middle of the table: 8 different neurons
    fired = 0
    for q in range(60):
        sb, so = b.run_pass(224), o.run_pass(224)
        assert_same_stats(sb, so, f"pass {q}")
        fired += so.fired
    assert fired > 300
    assert_same_state(b, o)
    sb, so = b.prune_and_grow(), o.prune_and_grow()
    assert (sb.pruned, sb.appended, sb.n_after) == (so.pruned, so.appended, so.n_after) and so.appended > 20 and so.pruned > 0
    assert_same_state(b, o, "after the structural step")
    for q in range(10):
        assert_same_stats(b.run_pass(224), o.run_pass(224), f"pass {q} after the structural step")
    assert_same_state(b, o, "end")


def test_ragged_table_and_passes_visits_exact():
    """Table length not a multiple of the line or the 16-record block (sample_block 1, 8, 16), passes of 1, 7, 223, 257, 4099 events, per-event
    ticks that need the ANCIENT sentinel (clock far beyond 2^30 past the last fires): lastVisited and the candidate
    count are order-free and must equal the oracle's exactly; single-warp passes are bit-exact as a whole."""
    rng = np.random.default_rng(8)
    N, n = 300, 4099
    syn = random_graph(rng, n, N, 0.4, 1.0, dst_lo=8)
    for block in (1, 8, 16):
        b, o = pair(n_input=4, n_output=4, n_hidden=N - 8, n_syn=n, sample_block=block, window_pre=10**12, refractory=1000)
        pre = rng.integers(1, 50, N).astype(np.uint64)
        for x in (b, o):
            x.upload_synapses(syn); x.upload_timestamps(pre, None)
            x.clock = 3 * 2**30 + 17; x.set_reward(0.1)
        for events in (1, 7, 223):
            assert_same_stats(b.run_pass(events), o.run_pass(events), f"{events} events")
        assert_same_state(b, o, f"block {block}, single-warp passes")
        for events in (257, 4099):
            sb, so = b.run_pass(events), o.run_pass(events)
            assert (sb.events, sb.candidates) == (so.events, so.candidates)
            assert np.array_equal(b.timestamps()[1], o.timestamps()[1])
            assert so.gated > 0 and 0.5 * so.gated <= sb.gated <= 1.5 * so.gated + 20      # 300 neurons, 8 warps unordered: loose


def test_future_source_timestamps_take_the_exact_gate():
    """lastFired values of SOURCE neurons uploaded in the future of the clock (gate word SLACK_EXACT): the kernel takes
    the exact 64-bit window test for them; single-warp passes stay bit-exact against the oracle. (Future timestamps of
    DESTINATION neurons are outside the comparison on purpose: PARALLEL execution treats a fire from an event's own
    future symmetrically — DESIGN.md §2 — where the serial order never sees one.)"""
    rng = np.random.default_rng(2)
    N, n = 64, 48 * 64
    syn = np.zeros(n, O.SYN_DTYPE)
    syn["src"] = rng.integers(0, N, n); syn["dst"] = 16 + np.repeat(np.arange(48), 64)
    syn["w"] = rng.uniform(0.3, 1.0, n).astype(np.float32)
    lf = rng.integers(1, 900, N).astype(np.uint64)
    lf[:16:2] = 1000 + rng.integers(1, 4000, 8)                   # sources a few ticks ahead of the clock
    lf[3] = 1000 + 2**31                                           # far ahead
    b, o = pair(n_input=8, n_output=8, n_hidden=N - 16, n_syn=n, sample_block=8, window_pre=5000, refractory=400)
    for x in (b, o):
        x.upload_synapses(syn); x.upload_timestamps(lf, None); x.clock = 1000; x.set_reward(0.0)
    for p in range(30):
        assert_same_stats(b.run_pass(224), o.run_pass(224), f"pass {p}")
    assert_same_state(b, o)


def test_statistical_parity_and_visits_at_2m_synapses():
    """Many warps in flight (2M synapses, 700k-event passes, sample_block 1, 8 and 16): lastVisited and the first pass's
    candidates equal the oracle's exactly; gated / fired counts within 2 % + 5 sigma (refractory period = 0.7 pass)."""
    for block in (1, 8, 16):
        b, o = pair(n_input=64, n_output=64, n_hidden=200_000, n_syn=2_000_003, sample_block=block,
                    window_pre=3_000_000, refractory=500_000)
        b.init_graph(capi.GRAPH_ER_BETA, 5); o.init_graph(capi.GRAPH_ER_BETA, 5)
        pre = np.random.default_rng(9).integers(1, 1_000_000, 200_128).astype(np.uint64)
        for x in (b, o):
            x.upload_timestamps(pre, None); x.clock = 1_000_000; x.set_reward(0.0)
        for p in range(3):
            sb, so = b.run_pass(700_001), o.run_pass(700_001)
            assert sb.events == so.events and (p > 0 or sb.candidates == so.candidates)
            for f in ("gated", "fired"):
                g, w = getattr(sb, f), getattr(so, f)
                assert abs(g - w) <= 0.02 * w + 5 * np.sqrt(w + 1), (block, p, f, g, w)
            assert np.array_equal(b.timestamps()[1], o.timestamps()[1]), f"lastVisited differs at pass {p}"
        b.close()


def test_parallel_budget_is_respected_and_close_to_the_oracle():
    """PARALLEL execution with the saturating spike budget (brain.metal:85-98, brain.cpp:90; runs on the 64-bit line
    kernel's turn-based path): never more than max_spikes_per_pass fires in a pass, the budget binds (the unbudgeted
    run fires more), and gated counts stay within 15 % of the serial oracle, which stops gating at the same budget
    (where in the pass the budget runs out is fuzzy by the ~10 % of a pass that is in flight unordered)."""
    over = dict(n_input=64, n_output=64, n_hidden=50_000, n_syn=1_000_003, sample_block=8, window_pre=3_000_000,
                refractory=200_000, max_spikes_per_pass=1500)
    b, o = pair(**over)
    b.init_graph(capi.GRAPH_ER_BETA, 5); o.init_graph(capi.GRAPH_ER_BETA, 5)
    pre = np.random.default_rng(9).integers(1, 1_000_000, 50_128).astype(np.uint64)
    for x in (b, o):
        x.upload_timestamps(pre, None); x.clock = 1_000_000; x.set_reward(0.0)
    for p in range(3):
        sb, so = b.run_pass(500_000), o.run_pass(500_000)
        assert sb.fired <= 1500 and so.fired <= 1500
        assert so.fired == 1500 and sb.fired >= 1450, (sb.fired, so.fired)      # the budget binds in both
        assert abs(sb.gated - so.gated) <= 0.15 * so.gated + 50, (sb.gated, so.gated)
        assert np.array_equal(b.timestamps()[1], o.timestamps()[1])
