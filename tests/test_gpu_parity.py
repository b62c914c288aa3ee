"""GPU parity tests: the CUDA path (through the C-ABI) against the oracle on the same seeded inputs.

Bar (BASELINE.json north_star): SERIAL / conflict-free execution bit-exact on fire decisions,
timestamps, prune/append order; weights here are also required bit-exact (tolerance 0, tighter
than the 1e-6 relative the north star allows). Fully PARALLEL execution: statistical bounds,
stated per test.
"""
import numpy as np
import pytest

from abnn_b200 import Brain, BrainEngine, FunctionalDataset, capi
from oracle import pyoracle as O
from tests.helpers import assert_same_state, assert_same_stats, random_graph
from tests.test_oracle import kat_graph

pytestmark = pytest.mark.gpu

TOY = dict(n_input=256, n_output=256, n_hidden=10_000, n_syn=1_000_000)   # BASELINE.json configs[0]


def pair(profile, **over):
    p = O.default_params(profile, **over)
    return Brain(p), O.OracleB(p)


# ---- metal-parity profile: the reference kernel itself -------------------------------------------
def test_metal_parity_kat_vs_verbatim_reference_kernel():
    """GPU SERIAL metal-parity == brain.metal compiled verbatim (Oracle A), SURVEY.md §8c KAT."""
    syn = kat_graph()
    b, o = pair(capi.PROFILE_METAL_PARITY, n_input=16, n_output=16, n_hidden=32, n_syn=1024)
    b.upload_synapses(syn); o.upload_synapses(syn)
    b.set_reward(0.1); o.set_reward(0.1)
    a = O.OracleA(syn, 64, reward=0.1, hold_clock=True) if O.have_ref() else None
    for p in range(8):
        sb, so = b.run_pass(1024), o.run_pass(1024)
        assert_same_stats(sb, so, f"pass {p}")
        assert_same_state(b, o, f"pass {p}")
        if a is not None:
            a.run_pass(1024)
            assert b.download_synapses().tobytes() == a.syn.tobytes()
            assert np.array_equal(b.timestamps()[0], a.lastF.astype(np.uint64))
            assert 2560 - sb.fired == a.st.budget


@pytest.mark.parametrize("seed", [1, 2])
def test_metal_parity_random_graph(seed):
    rng = np.random.default_rng(seed)
    N, n = 512, 256 * 40
    syn = random_graph(rng, n, N)
    pre = rng.integers(1, 12, N).astype(np.uint64)
    b, o = pair(capi.PROFILE_METAL_PARITY, n_input=16, n_output=16, n_hidden=N - 32, n_syn=n)
    for x in (b, o):
        x.upload_synapses(syn); x.upload_timestamps(pre, None); x.clock = int(pre.max()) + 1; x.set_reward(0.2 * seed - 0.3)
    for p in range(10):
        assert_same_stats(b.run_pass(n), o.run_pass(n), f"pass {p}")
        assert_same_state(b, o, f"pass {p}")


# ---- north-star profile, SERIAL execution: bit-exact ------------------------------------------------
def test_north_star_serial_toy_engine_loop():
    """configs[0] shape (256/256/10k, 1M synapses): reference graph, sine input, teacher forcing, reward,
    read-out; events per pass reduced to 150k so the serial walk finishes in seconds."""
    over = dict(TOY, exec_mode=capi.EXEC_SERIAL, window_pre=400_000, refractory=50_000, seed=42, p_new=0.05,
                syn_capacity=1_000_500)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    b.build_random_graph(1); o.init_graph(capi.GRAPH_REFERENCE, 1)
    stim = FunctionalDataset()
    even = False
    total_fired = 0
    for p in range(6):
        vin, exp = stim.nextInput(), stim.nextExpected()
        for x in (b, o):
            x.inject_inputs(vin, 1000.0)
            x.teacher_force(exp, 1.0 if even else 0.0)
        even = not even
        sb, so = b.run_pass(150_000), o.run_pass(150_000)
        assert_same_stats(sb, so, f"pass {p}")
        total_fired += so.fired
        assert np.array_equal(b.read_outputs(), o.read_outputs())
        rb, ro = b.readout_filtered(exp), o.readout_filtered(exp)
        assert rb.tobytes() == ro.tobytes(), f"filtered read-out differs at pass {p}"
        if p == 2:
            b.set_reward(0.05); o.set_reward(0.05)
    assert_same_state(b, o, "after 6 passes")
    assert total_fired > 100, "workload degenerate: nothing fired"
    sb, so = b.prune_and_grow(), o.prune_and_grow()
    assert (sb.appended, sb.pruned, sb.n_after, sb.dropped) == (so.appended, so.pruned, so.n_after, so.dropped)
    assert so.appended == 500 and so.dropped > 0        # capacity reached: the tail of the ordered list is dropped
    assert_same_state(b, o, "after growth")


def test_north_star_serial_per_pass_clock_and_budget():
    rng = np.random.default_rng(5)
    N, n = 4096, 200_000
    syn = random_graph(rng, n, N, 0.5, 1.0)
    pre = rng.integers(1, 9, N).astype(np.uint64)
    over = dict(n_input=32, n_output=32, n_hidden=N - 64, n_syn=n, exec_mode=capi.EXEC_SERIAL,
                clock_mode=capi.CLOCK_PER_PASS, window_pre=5, refractory=2, max_spikes_per_pass=500)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    for x in (b, o):
        x.upload_synapses(syn); x.upload_timestamps(pre, None); x.clock = 9; x.set_reward(-0.1)
    for p in range(4):
        sb, so = b.run_pass(60_000), o.run_pass(60_000)
        assert_same_stats(sb, so, f"pass {p}")
        assert so.fired <= 500
    assert_same_state(b, o)


@pytest.mark.parametrize("block", [2, 8, 32])
def test_block_sampler_serial_bit_exact(block):
    """sample_block > 1 (one Philox draw per run of `block` records; 8 = one 128-byte HBM line): SERIAL
    execution bit-exact against the oracle, table length not a multiple of the block, growth on."""
    rng = np.random.default_rng(block)
    N, n = 3000, 100_003
    syn = random_graph(rng, n, N, 0.2, 1.0, dst_lo=16)
    pre = rng.integers(1, 40_000, N).astype(np.uint64)
    over = dict(n_input=16, n_output=16, n_hidden=N - 32, n_syn=n, exec_mode=capi.EXEC_SERIAL, sample_block=block,
                window_pre=60_000, refractory=30_000, p_new=0.1, syn_capacity=n + 5000)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    for x in (b, o):
        x.upload_synapses(syn); x.upload_timestamps(pre, None); x.clock = 40_000; x.set_reward(0.1)
    for p in range(3):
        sb, so = b.run_pass(50_001), o.run_pass(50_001)
        assert_same_stats(sb, so, f"pass {p}")
        assert so.gated > 500
    assert_same_state(b, o)
    sb, so = b.prune_and_grow(), o.prune_and_grow()
    assert (sb.appended, sb.n_after) == (so.appended, so.n_after) and so.appended > 0
    assert_same_state(b, o, "after growth")


@pytest.mark.parametrize("block", [1, 8])
def test_parallel_block_sampler_visits_exact_and_counts_close(block):
    """PARALLEL execution with the iid and the line-granular sampler: lastVisited is an order-free max, so
    it must equal the oracle's exactly (proves every event touched the same record at the same tick);
    gated / fired counts within 8 % + 5 sigma (the refractory period here is ~ one pass, the worst case
    for unordered execution: a fire only blocks the events that run after it landed)."""
    over = dict(n_input=64, n_output=64, n_hidden=200_000, n_syn=2_000_003, exec_mode=capi.EXEC_PARALLEL,
                sample_block=block, window_pre=3_000_000, refractory=500_000)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    b.init_graph(capi.GRAPH_ER_BETA, 5); o.init_graph(capi.GRAPH_ER_BETA, 5)
    rng = np.random.default_rng(9)
    pre = rng.integers(1, 1_000_000, 200_128).astype(np.uint64)
    for x in (b, o):
        x.upload_timestamps(pre, None); x.clock = 1_000_000; x.set_reward(0.0)
    for p in range(3):
        sb, so = b.run_pass(700_001), o.run_pass(700_001)
        assert sb.events == so.events and sb.candidates == so.candidates or p > 0
        for f in ("gated", "fired"):
            g, w = getattr(sb, f), getattr(so, f)
            assert abs(g - w) <= 0.08 * w + 5 * np.sqrt(w + 1), (p, f, g, w)
        assert np.array_equal(b.timestamps()[1], o.timestamps()[1]), f"lastVisited differs at pass {p}"


# ---- EXACT execution: parallel and bit-identical to the serial order -----------------------------------
@pytest.mark.parametrize("block,clock_mode", [(1, capi.CLOCK_PER_EVENT), (8, capi.CLOCK_PER_EVENT), (8, capi.CLOCK_PER_PASS)])
def test_exact_mode_bit_exact_toy_full_size(block, clock_mode):
    """configs[0] at FULL size (256/256/10k neurons, 1M synapses, 1M-event passes, reference graph, sine
    input, teacher forcing, reward, growth + pruning): conflict-free parallel execution must reproduce
    the serial oracle bit for bit — fire decisions, both timestamp arrays, weights, append order."""
    per_pass = clock_mode == capi.CLOCK_PER_PASS
    over = dict(TOY, exec_mode=capi.EXEC_EXACT, sample_block=block, clock_mode=clock_mode, seed=42,
                window_pre=5 if per_pass else 2_000_000, refractory=2 if per_pass else 100_000,
                p_new=0.02, w_prune=0.03, syn_capacity=1_050_000)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    b.build_random_graph(1); o.init_graph(capi.GRAPH_REFERENCE, 1)
    stim = FunctionalDataset()
    even = False
    fired = 0
    n_pass = 8 if per_pass else 5          # per-pass clock: nothing can gate before pass 3 (refractory vs lastFired = 0)
    for p in range(n_pass):
        vin, exp = stim.nextInput(), stim.nextExpected()
        for x in (b, o):
            x.inject_inputs(vin, 1000.0); x.teacher_force(exp, 1.0 if even else 0.0)
        even = not even
        sb, so = b.run_pass(1_000_000), o.run_pass(1_000_000)
        assert_same_stats(sb, so, f"pass {p}")
        fired += so.fired
        assert b.readout_filtered(exp).tobytes() == o.readout_filtered(exp).tobytes()
        if p == n_pass - 3:
            ssb, sso = b.prune_and_grow(), o.prune_and_grow()
            assert (ssb.pruned, ssb.appended, ssb.n_after) == (sso.pruned, sso.appended, sso.n_after)
            assert sso.appended > 0
            b.set_reward(0.02); o.set_reward(0.02)
    assert fired > 1000
    assert_same_state(b, o, "after all passes")


def test_exact_mode_er_graph_many_conflicts():
    """Small neuron count, many events per destination per pass (long per-destination chains)."""
    over = dict(n_input=32, n_output=32, n_hidden=2000, n_syn=300_000, exec_mode=capi.EXEC_EXACT, sample_block=8,
                window_pre=5_000_000, refractory=3_000, p_new=0.01, syn_capacity=400_000)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    b.init_graph(capi.GRAPH_ER_BETA, 3); o.init_graph(capi.GRAPH_ER_BETA, 3)
    rng = np.random.default_rng(1)
    pre = rng.integers(1, 100_000, 2064).astype(np.uint64)
    for x in (b, o):
        x.upload_timestamps(pre, None); x.clock = 100_000; x.set_reward(-0.5)
    for p in range(3):
        sb, so = b.run_pass(500_000), o.run_pass(500_000)
        assert_same_stats(sb, so, f"pass {p}")
        assert so.gated > 100_000 and so.fired > 1000
    assert_same_state(b, o)
    sb, so = b.prune_and_grow(), o.prune_and_grow()
    assert (sb.appended, sb.n_after) == (so.appended, so.n_after) and so.appended > 10
    assert_same_state(b, o, "after growth")


def test_exact_mode_tiny_network_thousands_of_events_per_destination():
    """24 neurons under 60,000-event passes with a refractory period of 2 ticks: every destination collects thousands of
    open events per pass (phase 3 sorts its bucket with the heap-sort path) and walks them all; block and iid sampler."""
    rng = np.random.default_rng(77)
    N, n = 24, 4096
    syn = random_graph(rng, n, N, 0.05, 1.0, dst_lo=4)
    for block in (1, 16):
        b, o = pair(capi.PROFILE_NORTH_STAR, n_input=4, n_output=4, n_hidden=N - 8, n_syn=n, exec_mode=capi.EXEC_EXACT,
                    sample_block=block, window_pre=10**9, refractory=2)
        for x in (b, o):
            x.upload_synapses(syn); x.upload_timestamps(np.full(N, 1, np.uint64), None); x.clock = 1000; x.set_reward(0.3)
        for p in range(2):
            sb, so = b.run_pass(60_000), o.run_pass(60_000)
            assert_same_stats(sb, so, f"block {block} pass {p}")
            assert so.gated > 30_000 and so.fired > 1000
        assert_same_state(b, o, f"block {block}")


# ---- dst-sorted table layout (ABNN_TABLE_DST_SORTED) ----------------------------------------------------
@pytest.mark.parametrize("mode", [capi.EXEC_SERIAL, capi.EXEC_EXACT])
def test_dst_sorted_table_bit_exact(mode):
    """The device's stable radix sort by dst == the oracle's counting sort, after upload and after growth;
    SERIAL and EXACT execution over the sorted table stay bit-exact."""
    rng = np.random.default_rng(21)
    N, n = 5000, 300_007
    syn = random_graph(rng, n, N, 0.2, 1.0, dst_lo=16)
    pre = rng.integers(1, 40_000, N).astype(np.uint64)
    over = dict(n_input=16, n_output=16, n_hidden=N - 32, n_syn=n, exec_mode=mode, sample_block=8,
                table_order=capi.TABLE_DST_SORTED, window_pre=250_000, refractory=30_000, p_new=0.1, w_prune=0.21,
                syn_capacity=n + 20_000)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    for x in (b, o):
        x.upload_synapses(syn); x.upload_timestamps(pre, None); x.clock = 40_000; x.set_reward(0.1)
    assert b.download_synapses().tobytes() == syn[np.argsort(syn["dst"], kind="stable")].tobytes()
    for p in range(3):
        sb, so = b.run_pass(100_001), o.run_pass(100_001)
        assert_same_stats(sb, so, f"pass {p}")
        assert so.gated > 500
        if p == 1:
            ssb, sso = b.prune_and_grow(), o.prune_and_grow()
            assert (ssb.pruned, ssb.appended, ssb.n_after) == (sso.pruned, sso.appended, sso.n_after)
            assert sso.appended > 0 and sso.pruned > 0
    assert_same_state(b, o)


@pytest.mark.parametrize("mode", [capi.EXEC_SERIAL, capi.EXEC_EXACT])
def test_dst_interleaved_table_bit_exact(mode):
    """ABNN_TABLE_DST_INTERLEAVED: the device's radix sort + per-group interleave == the oracle's, after upload and after a
    structural step that prunes and grows (the order is re-derived on both sides); SERIAL and EXACT stay bit-exact."""
    rng = np.random.default_rng(22)
    N, n = 5000, 300_007
    syn = random_graph(rng, n, N, 0.2, 1.0, dst_lo=16)
    pre = rng.integers(1, 40_000, N).astype(np.uint64)
    over = dict(n_input=16, n_output=16, n_hidden=N - 32, n_syn=n, exec_mode=mode, sample_block=8,
                table_order=capi.TABLE_DST_INTERLEAVED, window_pre=250_000, refractory=30_000, p_new=0.1, w_prune=0.21,
                syn_capacity=n + 20_000)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    for x in (b, o):
        x.upload_synapses(syn); x.upload_timestamps(pre, None); x.clock = 40_000; x.set_reward(0.1)
    got = b.download_synapses()
    assert got.tobytes() == o.download_synapses().tobytes()
    s = syn[np.argsort(syn["dst"], kind="stable")]                      # independent restatement of the order in numpy
    first = np.searchsorted(s["dst"], s["dst"], side="left")
    rank = np.arange(n) - first
    assert got.tobytes() == s[np.lexsort((s["dst"] & 7, rank, s["dst"] >> 3))].tobytes()
    for p in range(3):
        sb, so = b.run_pass(100_001), o.run_pass(100_001)
        assert_same_stats(sb, so, f"pass {p}")
        assert so.gated > 500
        if p == 1:
            ssb, sso = b.prune_and_grow(), o.prune_and_grow()
            assert (ssb.pruned, ssb.appended, ssb.n_after) == (sso.pruned, sso.appended, sso.n_after)
            assert sso.appended > 0 and sso.pruned > 0
            assert b.download_synapses().tobytes() == o.download_synapses().tobytes()
    assert_same_state(b, o)


def test_line_kernel_sorted_table_single_warp_chains_exact():
    """PARALLEL line kernel over a dst-sorted table with ONE chunk in flight per destination: 64 destinations
    with 4096 synapses each and passes of 256 events (= one warp, one chunk). Lines of a chunk share
    destinations, so the in-warp chain resolution (match.any + ballot) carries every same-destination
    dependency, and the result must equal the serial oracle bit for bit."""
    rng = np.random.default_rng(33)
    N, n = 64, 64 * 4096
    syn = np.zeros(n, O.SYN_DTYPE)
    syn["src"] = rng.integers(0, N, n); syn["dst"] = np.repeat(np.arange(N), 4096)
    syn["w"] = rng.uniform(0.3, 1.0, n).astype(np.float32)
    over = dict(n_input=8, n_output=8, n_hidden=N - 16, n_syn=n, exec_mode=capi.EXEC_PARALLEL, sample_block=8,
                table_order=capi.TABLE_DST_SORTED, window_pre=10**9, refractory=3)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    for x in (b, o):
        x.upload_synapses(syn); x.upload_timestamps(np.full(N, 1, np.uint64), None); x.clock = 1000; x.set_reward(0.2)
    fired = 0
    for p in range(40):
        sb, so = b.run_pass(256), o.run_pass(256)
        assert_same_stats(sb, so, f"pass {p}")
        fired += so.fired
    assert fired > 500
    assert_same_state(b, o)


def test_line_kernel_sorted_table_statistical():
    """PARALLEL line kernel, dst-sorted ER graph at 2M synapses: lastVisited exact (order-free max, one RED per
    run of equal destinations), gated / fired counts within 8 % + 5 sigma of the serial oracle."""
    over = dict(n_input=64, n_output=64, n_hidden=200_000, n_syn=2_000_003, exec_mode=capi.EXEC_PARALLEL,
                sample_block=8, table_order=capi.TABLE_DST_SORTED, window_pre=3_000_000, refractory=500_000)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    b.init_graph(capi.GRAPH_ER_BETA, 5); o.init_graph(capi.GRAPH_ER_BETA, 5)
    assert b.download_synapses().tobytes() == o.download_synapses().tobytes()
    rng = np.random.default_rng(9)
    pre = rng.integers(1, 1_000_000, 200_128).astype(np.uint64)
    for x in (b, o):
        x.upload_timestamps(pre, None); x.clock = 1_000_000; x.set_reward(0.0)
    for p in range(3):
        sb, so = b.run_pass(700_001), o.run_pass(700_001)
        assert sb.events == so.events and (p > 0 or sb.candidates == so.candidates)
        for f in ("gated", "fired"):
            g, w = getattr(sb, f), getattr(so, f)
            assert abs(g - w) <= 0.08 * w + 5 * np.sqrt(w + 1), (p, f, g, w)
        assert np.array_equal(b.timestamps()[1], o.timestamps()[1]), f"lastVisited differs at pass {p}"


# ---- PARALLEL execution ------------------------------------------------------------------------------
def test_parallel_conflict_free_is_bit_exact():
    """When no two events of a pass share a destination, PARALLEL == SERIAL order bit for bit:
    SWEEP over a table whose dst values are a permutation (each dst once), unlimited budget."""
    rng = np.random.default_rng(11)
    N = 1 << 16
    syn = np.zeros(N, O.SYN_DTYPE)
    syn["src"] = rng.integers(0, N, N)
    syn["dst"] = rng.permutation(N)
    syn["w"] = rng.uniform(0.05, 1.0, N).astype(np.float32)
    pre = rng.integers(1, 30_000, N).astype(np.uint64)
    for clock_mode, win, refr in ((capi.CLOCK_PER_EVENT, 60_000, 20_000), (capi.CLOCK_PER_PASS, 5, 2)):
        over = dict(n_input=64, n_output=64, n_hidden=N - 128, n_syn=N, sampler=capi.SAMPLER_SWEEP,
                    exec_mode=capi.EXEC_PARALLEL, clock_mode=clock_mode, window_pre=win, refractory=refr)
        b, o = pair(capi.PROFILE_NORTH_STAR, **over)
        pre_c = pre if clock_mode == capi.CLOCK_PER_EVENT else (pre % 7 + 1)
        for x in (b, o):
            x.upload_synapses(syn); x.upload_timestamps(pre_c, None); x.clock = int(pre_c.max()) + 1; x.set_reward(0.3)
        for p in range(3):
            sb, so = b.run_pass(N), o.run_pass(N)
            assert_same_stats(sb, so, f"mode {clock_mode} pass {p}")
            assert so.gated > (1000 if p == 0 else 0)
        assert_same_state(b, o, f"clock mode {clock_mode}")


def test_engine_step_graph_replay_is_bit_exact():
    """abnn_engine_step (stage frame + inject + teacher forcing + pass + read-out as one call, recorded into a CUDA
    graph once the state repeats — the third call — and replayed afterwards) on a conflict-free PARALLEL workload: every pass equals
    the oracle driven by the separate calls, bit for bit, and the replay really happened."""
    rng = np.random.default_rng(12)
    N = 1 << 15
    syn = np.zeros(N, O.SYN_DTYPE)
    syn["src"] = rng.integers(0, N, N); syn["dst"] = rng.permutation(N)
    syn["w"] = rng.uniform(0.05, 1.0, N).astype(np.float32)
    pre = rng.integers(1, 30_000, N).astype(np.uint64)
    over = dict(n_input=64, n_output=64, n_hidden=N - 128, n_syn=N, sampler=capi.SAMPLER_SWEEP, exec_mode=capi.EXEC_PARALLEL,
                window_pre=60_000, refractory=20_000, reward_window=3)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    for x in (b, o):
        x.upload_synapses(syn); x.upload_timestamps(pre, None); x.clock = 30_001; x.set_reward(0.3)
    stim = FunctionalDataset(64, 64)
    for p in range(8):
        vin, exp = stim.nextInput(), stim.nextExpected()
        rb = b.engine_step(vin, exp, 1000.0, float(p & 1), N, want_rates=True)
        o.inject_inputs(vin, 1000.0); o.teacher_force(exp, float(p & 1)); o.run_pass(N)
        assert rb.tobytes() == o.readout_filtered(exp).tobytes(), f"pass {p}"
    assert_same_state(b, o, "after 8 engine steps")
    assert b.get_loss() == o.get_loss() and o.get_loss()[1] == 2


def test_engine_step_exact_mode_is_captured_and_bit_exact():
    """EXACT execution keeps its event counts on the device (buffers sized by the events of the pass), so abnn_engine_step records it into a CUDA
    graph like a PARALLEL pass: line sampler over the interleaved table, conflicts and growth on, 8 engine steps — every
    filtered read-out and the final state equal the oracle bit for bit."""
    over = dict(n_input=64, n_output=64, n_hidden=3000, n_syn=120_000, exec_mode=capi.EXEC_EXACT, sample_block=8,
                table_order=capi.TABLE_DST_INTERLEAVED, window_pre=400_000, refractory=20_000, p_new=0.05, reward_window=3,
                syn_capacity=130_000)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    b.build_random_graph(1); o.init_graph(capi.GRAPH_REFERENCE, 1)
    stim = FunctionalDataset(64, 64)
    for p in range(8):
        vin, exp = stim.nextInput(), stim.nextExpected()
        rb = b.engine_step(vin, exp, 1000.0, float(p & 1), 100_000, want_rates=True)
        o.inject_inputs(vin, 1000.0); o.teacher_force(exp, float(p & 1)); so = o.run_pass(100_000)
        assert rb.tobytes() == o.readout_filtered(exp).tobytes(), f"pass {p}"
    assert so.gated > 1000
    assert_same_state(b, o, "after 8 EXACT engine steps")
    assert b.get_loss() == o.get_loss() and o.get_loss()[1] == 2


def test_parallel_statistical_parity_toy():
    """configs[0] at full size (1M synapses, 1M-event passes), fully parallel. Bounds: gated and fired
    counts within 6 % of the oracle's (+ 5 sigma Poisson; the library keeps at most 1/16 of a pass in
    flight, which is the granularity at which unordered execution follows event order), mean weight within 1e-4 absolute, weight
    histogram L1 distance below 1 %, lastVisited identical (order-free max)."""
    over = dict(TOY, exec_mode=capi.EXEC_PARALLEL, window_pre=2_000_000, refractory=100_000, seed=42)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    b.build_random_graph(1); o.init_graph(capi.GRAPH_REFERENCE, 1)
    assert b.download_synapses().tobytes() == o.download_synapses().tobytes()
    stim = FunctionalDataset()
    even = False
    for p in range(4):
        vin, exp = stim.nextInput(), stim.nextExpected()
        for x in (b, o):
            x.inject_inputs(vin, 1000.0); x.teacher_force(exp, 1.0 if even else 0.0)
        even = not even
        sb, so = b.run_pass(1_000_000), o.run_pass(1_000_000)
        assert sb.events == so.events
        for f in ("gated", "fired"):
            g, w = getattr(sb, f), getattr(so, f)
            assert abs(g - w) <= 0.06 * w + 5 * np.sqrt(w + 1), (p, f, g, w)
    wb, wo = b.download_synapses()["w"], o.download_synapses()["w"]
    assert abs(float(wb.mean()) - float(wo.mean())) < 2e-4
    hb, _ = np.histogram(wb, bins=64, range=(0, 1)); ho, _ = np.histogram(wo, bins=64, range=(0, 1))
    assert np.abs(hb - ho).sum() / len(wo) < 0.01
    assert np.array_equal(b.timestamps()[1], o.timestamps()[1])


def test_parallel_deviation_across_seeds():
    """SURVEY.md §8c asks for PARALLEL bounds relative to the oracle's own run-to-run spread. configs[0] shape, line
    sampler, dst-sorted table, 8 Philox seeds, 3 passes each. Measured here: the oracle's seed-to-seed spread of the
    fire and gated counts is 0.2 %; PARALLEL sits about 1 % above the oracle for EVERY seed — a bias, not noise: the
    line kernel keeps at most a quarter of the refractory period in flight unordered (24k of 100k ticks here), and an
    event can still run before the fire of the same neuron that precedes it in event order (with 63k ticks in flight
    the bias was 3.7 %). At the benchmark shape the in-flight window is 1.2M ticks against a refractory period of
    300M (test_full_size_parallel_vs_exact_properties: below 1 %). Bounds asserted: every
    seed within 2 %, and the deviations of the 8 seeds within 0.5 % of each other (a property of the schedule, not of
    the seed)."""
    fired_o, fired_b, gated_o, gated_b = [], [], [], []
    for seed in range(100, 108):
        over = dict(TOY, exec_mode=capi.EXEC_PARALLEL, sample_block=8, table_order=capi.TABLE_DST_SORTED,
                    window_pre=2_000_000, refractory=100_000, seed=seed)
        b, o = pair(capi.PROFILE_NORTH_STAR, **over)
        b.init_graph(capi.GRAPH_ER_BETA, 3); o.init_graph(capi.GRAPH_ER_BETA, 3)
        pre = np.random.default_rng(5).integers(1, 1_000_000, 10_512).astype(np.uint64)
        fb = fo = gb = go = 0
        for x in (b, o):
            x.upload_timestamps(pre, None); x.clock = 1_000_000; x.set_reward(0.05)
        for p in range(3):
            sb, so = b.run_pass(1_000_000), o.run_pass(1_000_000)
            fb += sb.fired; fo += so.fired; gb += sb.gated; go += so.gated
        fired_o.append(fo); fired_b.append(fb); gated_o.append(go); gated_b.append(gb)
        b.close()
    for name, ob, bb in (("fired", fired_o, fired_b), ("gated", gated_o, gated_b)):
        ob, bb = np.array(ob, float), np.array(bb, float)
        dev = (bb - ob) / ob
        spread = ob.std(ddof=1) / ob.mean()
        assert ob.min() > 1000 and spread < 0.01
        assert np.abs(dev).max() <= 0.02, (name, dev, spread)
        assert dev.max() - dev.min() <= 0.005, (name, dev)


# ---- graph init / I/O ------------------------------------------------------------------------------
def test_er_beta_init_bit_exact_and_beta_moments():
    over = dict(n_input=64, n_output=64, n_hidden=100_000, n_syn=500_000)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    b.init_graph(capi.GRAPH_ER_BETA, 99); o.init_graph(capi.GRAPH_ER_BETA, 99)
    s = b.download_synapses()
    assert s.tobytes() == o.download_synapses().tobytes()
    assert abs(float(s["w"].mean()) - 0.2) < 2e-3            # Beta(2,8): mean 0.2, var 16/(100*11)
    assert abs(float(s["w"].var()) - 16 / 1100) < 5e-4


def test_bnn_roundtrip_and_shape_error(tmp_path):
    over = dict(n_input=8, n_output=8, n_hidden=500, n_syn=5000)
    p = O.default_params(capi.PROFILE_NORTH_STAR, **over)
    path = str(tmp_path / "model.bnn")
    with Brain(p) as b:
        b.build_random_graph(1)
        want = b.download_synapses()
        b.save(path)
    raw = open(path, "rb").read()
    assert len(raw) == 8 + 16 * 5000 and np.frombuffer(raw[:8], "<u4").tolist() == [5000, 516]   # brain.cpp:161-167
    with Brain(p) as b:
        b.load(path)
        assert b.download_synapses().tobytes() == want.tobytes()
    q = O.default_params(capi.PROFILE_NORTH_STAR, **dict(over, n_hidden=501))
    with Brain(q) as b:
        with pytest.raises(capi.AbnnError) as e:
            b.load(path)
        assert e.value.status == capi.ERR_SHAPE                                                  # brain.cpp:174
    # a pruned table (live count below n_syn) written by abnn_save_bnn loads back
    pr = O.default_params(capi.PROFILE_NORTH_STAR, **dict(over, w_prune=0.15))
    with Brain(pr) as b:
        b.build_random_graph(1)
        st = b.prune_and_grow()
        assert 0 < st.n_after < 5000
        kept = b.download_synapses()
        b.save(path)
    with Brain(pr) as b:
        b.load(path)
        assert b.download_synapses().tobytes() == kept.tobytes()


def test_bnn_v2_exact_resume(tmp_path):
    """.bnn v2: stop after 3 passes, reload into a fresh handle, continue — the continuation equals the
    uninterrupted run (and the oracle) bit for bit, including read-out filter state, reward window and
    staged growth candidates."""
    over = dict(n_input=64, n_output=64, n_hidden=3000, n_syn=120_000, exec_mode=capi.EXEC_EXACT, sample_block=8,
                table_order=capi.TABLE_DST_SORTED, window_pre=400_000, refractory=20_000, p_new=0.05, reward_window=4,
                syn_capacity=130_000)
    path = str(tmp_path / "state.bnn2")
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    b.build_random_graph(1); o.init_graph(capi.GRAPH_REFERENCE, 1)
    stim_b, stim_o = FunctionalDataset(64, 64), FunctionalDataset(64, 64)

    def step(x, stim, p):
        vin, exp = stim.nextInput(), stim.nextExpected()
        x.inject_inputs(vin, 1000.0); x.teacher_force(exp, float(p & 1))
        st = x.run_pass(100_000)
        return st, x.readout_filtered(exp)

    for p in range(3):
        sb, rb = step(b, stim_b, p); so, ro = step(o, stim_o, p)
        assert_same_stats(sb, so, f"pass {p}") ; assert rb.tobytes() == ro.tobytes()
    b.save_state(path)
    b.close()
    b2 = Brain(O.default_params(capi.PROFILE_NORTH_STAR, **over))
    b2.load_state(path)
    for p in range(3, 7):
        sb, rb = step(b2, stim_b, p); so, ro = step(o, stim_o, p)
        assert_same_stats(sb, so, f"resumed pass {p}"); assert rb.tobytes() == ro.tobytes()
        if p == 4:
            ssb, sso = b2.prune_and_grow(), o.prune_and_grow()
            assert (ssb.appended, ssb.n_after) == (sso.appended, sso.n_after) and sso.appended > 0
    assert_same_state(b2, o, "after resume")
    assert b2.get_loss() == o.get_loss() and o.get_loss()[1] == 1
    q = O.default_params(capi.PROFILE_NORTH_STAR, **dict(over, n_hidden=3001))
    with Brain(q) as other:
        with pytest.raises(capi.AbnnError) as e:
            other.load_state(path)
        assert e.value.status == capi.ERR_SHAPE
    # a state file means what the fields it was written under say: another table order / sampler / seed / window is refused
    # (a GIVEN-order table in a sorted handle would break the sorted insertion), while the execution mode may change
    for bad in (dict(table_order=capi.TABLE_AS_GIVEN), dict(sample_block=1), dict(seed=43), dict(window_pre=400_001)):
        with Brain(O.default_params(capi.PROFILE_NORTH_STAR, **dict(over, **bad))) as other:
            with pytest.raises(capi.AbnnError) as e:
                other.load_state(path)
            assert e.value.status == capi.ERR_SHAPE, bad
    with Brain(O.default_params(capi.PROFILE_NORTH_STAR, **dict(over, exec_mode=capi.EXEC_SERIAL))) as other:
        other.load_state(path)                                       # same semantics, other execution mode: accepted
        assert other.info().n_syn_local > 0
    short = str(tmp_path / "short.bnn2")
    raw = open(path, "rb").read()
    open(short, "wb").write(raw[:len(raw) - 1000])
    with Brain(O.default_params(capi.PROFILE_NORTH_STAR, **over)) as other:
        other.build_random_graph(1)
        before = other.download_synapses().tobytes()
        with pytest.raises(capi.AbnnError) as e:
            other.load_state(short)
        assert e.value.status == capi.ERR_IO
        assert other.download_synapses().tobytes() == before          # a truncated file leaves the handle untouched


# ---- structural plasticity ---------------------------------------------------------------------------
def test_prune_compaction_is_stable_and_matches_oracle():
    rng = np.random.default_rng(3)
    n, N = 3_000_001, 50_000
    syn = random_graph(rng, n, N, 0.0, 0.2)
    over = dict(n_input=16, n_output=16, n_hidden=N - 32, n_syn=n, w_prune=0.05)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    b.upload_synapses(syn); o.upload_synapses(syn)
    sb, so = b.prune_and_grow(), o.prune_and_grow()
    assert (sb.n_before, sb.pruned, sb.n_after) == (so.n_before, so.pruned, so.n_after)
    got = b.download_synapses()
    assert got.tobytes() == syn[syn["w"] >= np.float32(0.05)].tobytes()
    sb2 = b.prune_and_grow()                    # idempotent
    assert sb2.pruned == 0 and b.download_synapses().tobytes() == got.tobytes()


def test_prune_compaction_in_place_path_matches_oracle():
    """With room for a second table the prune goes out of place (count pass + scatter pass); params.prune_in_place = 1
    selects the in-place single-pass chained scan (k_compact), which is also what runs when the table takes more than a
    quarter of the device memory. Both must give the oracle's stable compaction."""
    rng = np.random.default_rng(5)
    n, N = 2_500_003, 40_000
    syn = random_graph(rng, n, N, 0.0, 0.2)
    b, o = pair(capi.PROFILE_NORTH_STAR, n_input=16, n_output=16, n_hidden=N - 32, n_syn=n, w_prune=0.07, prune_in_place=1)
    b.upload_synapses(syn); o.upload_synapses(syn)
    sb, so = b.prune_and_grow(), o.prune_and_grow()
    assert (sb.n_before, sb.pruned, sb.n_after) == (so.n_before, so.pruned, so.n_after) and so.pruned > 100000
    assert b.download_synapses().tobytes() == o.download_synapses().tobytes()


def test_structural_step_every_pass_fused_prune_merge_bit_exact():
    """BASELINE configs[4] regime: pruning + synaptogenesis after EVERY pass on a dst-sorted table. With room for every
    owned candidate the device removes the pruned records and inserts the new ones in ONE pass over the table
    (launch_prune_merge_sorted); the table, the counts and everything downstream must equal the oracle's sequential
    prune -> ordered insert bit for bit, including a pass with nothing to prune and one with nothing to grow."""
    rng = np.random.default_rng(77)
    N, n = 20_000, 800_003
    syn = random_graph(rng, n, N, 0.15, 1.0, dst_lo=32)
    pre = rng.integers(1, 60_000, N).astype(np.uint64)
    over = dict(n_input=32, n_output=32, n_hidden=N - 64, n_syn=n, exec_mode=capi.EXEC_EXACT, sample_block=8,
                table_order=capi.TABLE_DST_SORTED, window_pre=400_000, refractory=20_000, p_new=0.2, w_prune=0.16,
                w_init=0.17, syn_capacity=n + 200_000)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    for x in (b, o):
        x.upload_synapses(syn); x.upload_timestamps(pre, None); x.clock = 60_000; x.set_reward(0.1)
    grown = pruned = 0
    for p in range(5):
        sb, so = b.run_pass(200_003), o.run_pass(200_003)
        assert_same_stats(sb, so, f"pass {p}")
        ssb, sso = b.prune_and_grow(), o.prune_and_grow()
        assert (ssb.n_before, ssb.pruned, ssb.appended, ssb.dropped, ssb.n_after) == \
               (sso.n_before, sso.pruned, sso.appended, sso.dropped, sso.n_after), f"structural step {p}"
        assert sso.dropped == 0 and sso.n_after <= n + 200_000       # every candidate fitted: the fused path ran
        grown += sso.appended; pruned += sso.pruned
        assert b.download_synapses().tobytes() == o.download_synapses().tobytes(), f"table after structural step {p}"
    assert grown > 1000 and pruned > 1000
    sb, so = b.prune_and_grow(), o.prune_and_grow()                   # nothing staged: prune-only path, in place
    assert (sb.pruned, sb.appended) == (so.pruned, so.appended) and so.appended == 0
    assert_same_state(b, o)


def test_engine_loop_matches_oracle_per_pass_clock():
    """BrainEngine.run_one_pass order of operations (brain-engine.cpp:108-190) in the reference's own
    PER_PASS clock with SERIAL execution, 30 passes: filtered read-out bit-exact every pass."""
    over = dict(n_input=64, n_output=64, n_hidden=2000, n_syn=60_000, exec_mode=capi.EXEC_SERIAL,
                clock_mode=capi.CLOCK_PER_PASS, window_pre=5, refractory=2, max_spikes_per_pass=2560,
                reward_window=10)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    b.build_random_graph(1); o.init_graph(capi.GRAPH_REFERENCE, 1)
    eng = BrainEngine(b, events_per_pass=60_000, stimulus=FunctionalDataset(64, 64))
    stim_o = O.Dataset(64, 64)
    even = False
    for p in range(30):
        rb = eng.run_one_pass()
        vin, exp = stim_o.next_input(), stim_o.next_expected()
        o.inject_inputs(vin, 1000.0); o.teacher_force(exp, 1.0 if even else 0.0); even = not even
        o.run_pass(60_000)
        ro = o.readout_filtered(exp)
        assert rb.tobytes() == ro.tobytes(), f"pass {p}"
    assert b.get_loss() == o.get_loss() and o.get_loss()[1] == 3
    assert_same_state(b, o)


# ---- edge cases: empty and ragged inputs, error behaviour ------------------------------------------------------
@pytest.mark.parametrize("mode", [capi.EXEC_SERIAL, capi.EXEC_EXACT, capi.EXEC_PARALLEL])
def test_ragged_and_empty_passes(mode):
    """Table length not a multiple of the line (13 records), passes of 0, 1, 7, 255, 257 and 1000 events: counts,
    lastVisited and (SERIAL/EXACT) the whole state match the oracle; an empty table and a zero-event pass are no-ops."""
    rng = np.random.default_rng(8)
    N, n = 40, 13
    syn = random_graph(rng, n, N, 0.5, 1.0)
    over = dict(n_input=4, n_output=4, n_hidden=N - 8, n_syn=n, exec_mode=mode, sample_block=8, table_order=capi.TABLE_DST_SORTED,
                window_pre=10**6, refractory=50)
    b, o = pair(capi.PROFILE_NORTH_STAR, **over)
    for x in (b, o):
        x.upload_timestamps(np.full(N, 5, np.uint64), None); x.clock = 100
    for x in (b, o):                                   # empty table: every event is skipped
        st = x.run_pass(300)
        assert (st.gated, st.fired, st.candidates) == (0, 0, 0)
    b.upload_synapses(syn); o.upload_synapses(syn)
    total = 0
    for events in (0, 1, 7, 255, 257, 1000):
        sb, so = b.run_pass(events), o.run_pass(events)
        assert sb.events == so.events == events
        if mode == capi.EXEC_PARALLEL:
            assert sb.candidates == so.candidates or total > 0      # fires of earlier passes may differ
            assert np.array_equal(b.timestamps()[1], o.timestamps()[1])
        else:
            assert_same_stats(sb, so, f"{events} events")
        total += so.gated
    assert total > 100
    if mode != capi.EXEC_PARALLEL:
        assert_same_state(b, o)


def test_error_behaviour():
    """Bad arguments come back as status codes with a message; nothing throws across the ABI, the handle stays usable."""
    import ctypes as C
    p = O.default_params(capi.PROFILE_NORTH_STAR, n_input=8, n_output=8, n_hidden=100, n_syn=1000, syn_capacity=1000)
    with Brain(p) as b:
        with pytest.raises(capi.AbnnError) as e:
            b.inject_inputs(np.zeros(7, np.float32), 1000.0)                 # brain.cpp:75 expects n_input values
        assert e.value.status == capi.ERR_INVALID
        with pytest.raises(capi.AbnnError) as e:
            b.upload_synapses(np.zeros(1001, O.SYN_DTYPE))
        assert e.value.status == capi.ERR_CAPACITY
        with pytest.raises(capi.AbnnError) as e:
            b.load("/nonexistent/model.bnn")
        assert e.value.status == capi.ERR_IO
        # a record that names a neuron the handle does not have is refused (the kernels index per-neuron arrays with it),
        # from the host table and from a .bnn file alike; the handle is left with an empty table and stays usable
        for field in ("src", "dst"):
            t = np.zeros(10, O.SYN_DTYPE); t["w"] = 0.5; t[field][7] = 116                 # N = 116 neurons: 0 .. 115
            with pytest.raises(capi.AbnnError) as e:
                b.upload_synapses(t)
            assert e.value.status == capi.ERR_INVALID and b.info().n_syn_local == 0
        import struct, tempfile, os
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "bad.bnn")
            t = np.zeros(4, O.SYN_DTYPE); t["src"][2] = 0xFFFFFFFF
            with open(path, "wb") as f:
                f.write(struct.pack("<II", 4, 116)); f.write(t.tobytes())
            with pytest.raises(capi.AbnnError) as e:
                b.load(path)
            assert e.value.status == capi.ERR_INVALID
        b.build_random_graph(1)
        assert b.run_pass(500).events == 500
    for bad in (dict(sample_block=3), dict(exec_mode=7), dict(world_size=2, rank=2), dict(exec_mode=capi.EXEC_EXACT, src_view=capi.SRC_LIVE),
                dict(table_order=9)):
        q = O.default_params(capi.PROFILE_NORTH_STAR, n_input=8, n_output=8, n_hidden=100, n_syn=1000, **bad)
        h = C.c_void_p()
        assert capi.load().abnn_create(C.byref(q), C.byref(h)) in (capi.ERR_INVALID, capi.ERR_UNSUPPORTED), bad
        assert not h.value


# ---- full size (BASELINE.json configs[1]): size-independent properties --------------------------------------------------
@pytest.mark.parametrize("n_syn", [100_000_000, 1_000_000_000], ids=["configs1-100M", "configs2-1B"])
def test_full_size_parallel_vs_exact_properties(n_syn):
    """configs[1] / configs[2] — 5,000,000 hidden, 100,000,000 / 1,000,000,000 synapses, 150,000,000-event passes, the
    benchmark's layout (ABNN_PROFILE_B200: 256-byte sample groups over the interleaved table). The oracle would need minutes here, so the throughput kernel (PARALLEL) is checked against the EXACT mode
    (itself bit-identical to the oracle at the sizes above) through properties that do not depend on execution order:
    the same table after init + sort, the same pre-spike candidates in the first pass, identical lastVisited after every
    pass (an order-free max over the sampled events), and gated / fired counts and mean weight within 1 %."""
    import torch
    if torch.cuda.mem_get_info()[0] / 2**30 < 3.5 * 16 * n_syn / 2**30 + 12:
        pytest.skip("not enough free device memory")
    big = n_syn > 200_000_000                                   # 1B: no 16 GB table download, the order-free properties only
    events = 150_000_000
    over = dict(n_input=256, n_output=256, n_hidden=5_000_000, n_syn=n_syn, sample_block=16, table_order=capi.TABLE_DST_INTERLEAVED,
                window_pre=5 * events, refractory=2 * events, seed=42)
    n = 5_000_512
    rng = np.random.default_rng(7)
    lf = np.zeros(n, np.uint64)
    idx = rng.choice(n, size=n // 4, replace=False)
    lf[idx] = rng.integers(events, 6 * events, size=len(idx)).astype(np.uint64)
    res = {}
    for mode in (capi.EXEC_EXACT, capi.EXEC_PARALLEL):
        with Brain(O.default_params(capi.PROFILE_NORTH_STAR, exec_mode=mode, **over)) as b:
            b.init_graph(capi.GRAPH_ER_BETA, 1)
            b.upload_timestamps(lf, None); b.clock = 6 * events; b.set_reward(0.01)
            stats = [b.run_pass(events) for _ in range(2)]
            if big:
                res[mode] = (stats, b.timestamps()[1], 1.0, 0, 0)
            else:
                syn = b.download_synapses()
                assert np.all(np.diff((syn["dst"] >> 4).astype(np.int64)) >= 0)          # groups of 16 neurons in order
                res[mode] = (stats, b.timestamps()[1], float(syn["w"].mean(dtype=np.float64)),
                             int(syn["src"].astype(np.uint64).sum()), int(syn["dst"].astype(np.uint64).sum()))
    (se, lve, we, srce, dste), (sp, lvp, wp, srcp, dstp) = res[capi.EXEC_EXACT], res[capi.EXEC_PARALLEL]
    assert (srce, dste) == (srcp, dstp)                                  # same graph
    assert se[0].candidates == sp[0].candidates > 10_000_000             # same events, same gate decisions
    assert np.array_equal(lve, lvp), "lastVisited differs between EXACT and PARALLEL at full size"
    for a, b_ in zip(se, sp):
        assert a.events == b_.events == events
        assert abs(a.gated - b_.gated) <= 0.01 * a.gated and abs(a.fired - b_.fired) <= 0.01 * a.fired + 100, (a.gated, b_.gated, a.fired, b_.fired)
    assert abs(we - wp) < 1e-4 * we


def _host_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2**30
    except Exception:
        return 0.0


@pytest.mark.parametrize("n_syn", [100_000_000, 1_000_000_000], ids=["configs1-100M", "configs2-1B"])
def test_full_size_oracle_spot_check(n_syn):
    """BASELINE configs[1] / configs[2] shapes (5M hidden; 100M / 1B synapses; ABNN_PROFILE_B200 layout) against the ORACLE itself:
    the device's table is handed to Oracle B, both execute the same 2,000,000-event pass from the same warm state. EXACT
    execution must equal the oracle bit for bit (table, lastFired, lastVisited, counters) — this pins EXACT, which the
    full-size PARALLEL checks lean on, at the benchmark's size; PARALLEL (k_traverse_line32) must give the same candidates and
    lastVisited and gated / fired counts within 2 %. The 1B case needs about 70 GB of host memory (table + the oracle's copy
    and sort buffer) and is skipped on smaller hosts."""
    import torch
    need_dev = 3.5 * 16 * n_syn / 2**30 + 8
    if torch.cuda.mem_get_info()[0] / 2**30 < need_dev:
        pytest.skip(f"needs {need_dev:.0f} GB of free device memory")
    if _host_gb() < 5.0 * 16 * n_syn / 2**30 + 8:
        pytest.skip("not enough host memory for the oracle's copy of the table")
    import hashlib
    digest = lambda a: hashlib.sha256(np.ascontiguousarray(a)).hexdigest()
    events, full = 2_000_000, 150_000_000
    over = dict(n_input=256, n_output=256, n_hidden=5_000_000, n_syn=n_syn, window_pre=5 * full, refractory=2 * full, seed=42)
    n = 5_000_512
    rng = np.random.default_rng(7)
    lf = np.zeros(n, np.uint64)
    idx = rng.choice(n, size=n // 4, replace=False)
    lf[idx] = rng.integers(full, 6 * full, size=len(idx)).astype(np.uint64)
    res = {}
    table = None
    for mode in (capi.EXEC_EXACT, capi.EXEC_PARALLEL):
        with Brain(O.default_params(capi.PROFILE_B200, exec_mode=mode, **over)) as b:
            b.init_graph(capi.GRAPH_ER_BETA, 1)
            if table is None:
                table = b.download_synapses()                    # the pass-start table, for the oracle
            b.upload_timestamps(lf, None); b.clock = 6 * full; b.set_reward(0.01)
            st = b.run_pass(events)
            lfb, lvb = b.timestamps()
            after = b.download_synapses() if mode == capi.EXEC_EXACT else None
            res[mode] = (st, lfb, lvb, None if after is None else (digest(after["w"]), digest(after["dst"])))
            del after
    o = O.OracleB(O.default_params(capi.PROFILE_B200, exec_mode=capi.EXEC_SERIAL, **over))
    o.upload_synapses(table)
    del table
    assert o.n_syn_local() == n_syn
    o.upload_timestamps(lf, None); o.clock = 6 * full; o.set_reward(0.01)
    so = o.run_pass(events)
    olf, olv = o.timestamps()
    se, lfe, lve, syn_e = res[capi.EXEC_EXACT]
    assert_same_stats(se, so, "EXACT vs oracle")
    assert so.gated > 100_000 and so.fired > 1000
    assert np.array_equal(lfe, olf) and np.array_equal(lve, olv)
    osyn = o.download_synapses()
    assert syn_e == (digest(osyn["w"]), digest(osyn["dst"])), "EXACT table differs from the oracle's at full size"
    del osyn
    sp, lfp, lvp, _ = res[capi.EXEC_PARALLEL]
    assert (sp.events, sp.candidates) == (so.events, so.candidates)
    assert np.array_equal(lvp, olv), "PARALLEL lastVisited differs from the oracle at full size"
    assert abs(sp.gated - so.gated) <= 0.02 * so.gated and abs(sp.fired - so.fired) <= 0.02 * so.fired + 100, (sp.gated, so.gated, sp.fired, so.fired)
