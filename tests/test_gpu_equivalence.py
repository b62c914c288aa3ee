"""Sampler equivalence: does the line sampler (sample_block 8 / 16, one Philox draw per 128 / 256 bytes of the table)
drive the network like the iid sampler of SURVEY.md §8.0 (edge(e) = mulhi(philox(seed, e), N_SYN), README.md:77)?

Every synapse keeps the same sampling probability under any sample_block; what changes is the ARRIVAL of events at a
neuron. Over a DST_SORTED table the events of a sample group hit ONE neuron in a burst; once one of them fires, the rest
of the burst is refractory and wasted, which the iid sampler does not do — the fire and gated fractions come out a few
per cent lower (burst tail: about (B-1)/2 wasted events per fire). Over the DST_INTERLEAVED table (include/abnn.h) the
events of a group hit B different, adjacent neurons, each neuron's events arrive one at a time from independent draws —
the same point process as under iid sampling — and the dynamics agree within the seed-to-seed spread.

Measured per sampler over >= 8 seeds, engine loop of the reference (sine input, teacher forcing on alternate passes,
read-out, windowed loss -> reward), network in its active regime: gated fraction, fire fraction, weight histogram,
read-out loss. Bounds: 3 standard errors of the difference of the two 8-seed means (from the seed-to-seed spread of both
arms) plus a 0.5 % relative floor. The DST_SORTED line sampler is asserted to show its bias, so the test also documents
why the interleaved order exists.
"""
import numpy as np
import pytest

from abnn_b200 import Brain, FunctionalDataset, capi
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu

SEEDS = list(range(200, 208))
SAMPLERS = {   # name -> (sample_block, table_order)
    "iid": (1, capi.TABLE_AS_GIVEN),
    "line8_interleaved": (8, capi.TABLE_DST_INTERLEAVED),
    "line16_interleaved": (16, capi.TABLE_DST_INTERLEAVED),
    "line8_dst_sorted": (8, capi.TABLE_DST_SORTED),
}


def run_engine(shape, events, block, order, seed, exec_mode, passes, settle, reward_window=10):
    """`passes` passes of the reference's engine loop; returns gated and fire fraction over the passes after `settle`,
    the 64-bin weight histogram (fractions) and the mean of the windowed read-out losses."""
    n_neuron = shape["n_input"] + shape["n_output"] + shape["n_hidden"]
    p = O.default_params(capi.PROFILE_NORTH_STAR, exec_mode=exec_mode, sample_block=block, table_order=order, seed=seed,
                         window_pre=5 * events, refractory=2 * events, reward_window=reward_window, **shape)
    with Brain(p) as b:
        b.init_graph(capi.GRAPH_ER_BETA, 1)                       # the same graph for every sampler and seed
        rng = np.random.default_rng(7)
        lf = np.zeros(n_neuron, np.uint64)
        idx = rng.choice(n_neuron, size=n_neuron // 4, replace=False)
        lf[idx] = rng.integers(events, 6 * events, size=len(idx)).astype(np.uint64)
        b.upload_timestamps(lf, None); b.clock = 6 * events; b.set_reward(0.01)
        stim = FunctionalDataset(shape["n_input"], shape["n_output"])
        gated = fired = evs = 0
        losses, seen = [], 0
        for it in range(passes):
            vin, exp = stim.nextInput(), stim.nextExpected()
            b.inject_inputs(vin, 1000.0); b.teacher_force(exp, float(it & 1))
            st = b.run_pass(events)
            b.readout_step(exp)
            if it >= settle:
                gated += st.gated; fired += st.fired; evs += st.events
            loss, windows = b.get_loss()
            if windows > seen:
                losses.append(loss); seen = windows
        w = b.download_synapses()["w"]
        hist, _ = np.histogram(w, bins=64, range=(0.0, 1.0))
    return gated / evs, fired / evs, hist / len(w), float(np.mean(losses))


def compare(name, a, b, floor):
    """|mean(a) - mean(b)| within 3 standard errors of the difference + a relative floor."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    se = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
    diff = a.mean() - b.mean()
    bound = 3.0 * se + floor * abs(b.mean())
    return diff, bound, f"{name}: mean {a.mean():.6g} vs iid {b.mean():.6g} (diff {diff / b.mean():+.3%}, bound {bound / abs(b.mean()):.3%}, " \
                        f"seed spread {b.std(ddof=1) / abs(b.mean()):.3%})"


def check_equivalence(res, floor_counts=0.005):
    """res[name] = list over seeds of (gated fraction, fire fraction, weight histogram, mean loss)."""
    iid = res["iid"]
    report = []
    for name in ("line8_interleaved", "line16_interleaved"):
        for k, what, floor in ((0, "gated fraction", floor_counts), (1, "fire fraction", floor_counts), (3, "read-out loss", 0.02)):
            diff, bound, msg = compare(f"{name} {what}", [r[k] for r in res[name]], [r[k] for r in iid], floor)
            report.append(msg)
            assert abs(diff) <= bound, msg
        ha, hb = np.mean([r[2] for r in res[name]], axis=0), np.mean([r[2] for r in iid], axis=0)
        spread = np.mean([np.abs(r[2] - hb).sum() for r in iid])          # L1 distance of one iid seed from the iid mean
        l1 = np.abs(ha - hb).sum()
        report.append(f"{name} weight histogram L1 {l1:.5f} (iid seed-to-mean {spread:.5f})")
        assert l1 <= 3.0 * spread + 0.002, report[-1]
    # the dst-sorted line sampler: bursts of 8 events on one neuron -> fewer fires per event (documented bias)
    diff, bound, msg = compare("line8_dst_sorted fire fraction", [r[1] for r in res["line8_dst_sorted"]], [r[1] for r in iid], 0.005)
    report.append(msg)
    b_mean = np.mean([r[1] for r in iid])
    assert -0.10 * b_mean < diff < -0.005 * b_mean, msg
    return report


def test_sampler_equivalence_toy_exact_execution():
    """BASELINE configs[0] shape (256 / 256 / 10k neurons, 1M synapses, 1M-event passes), EXACT execution (no execution-
    order effects at all: what differs between the arms is the sampler and nothing else), 60 passes, 8 seeds."""
    shape = dict(n_input=256, n_output=256, n_hidden=10_000, n_syn=1_000_000)
    res = {name: [run_engine(shape, 1_000_000, blk, order, s, capi.EXEC_EXACT, passes=60, settle=20) for s in SEEDS]
           for name, (blk, order) in SAMPLERS.items()}
    for line in check_equivalence(res):
        print(line)


def test_sampler_equivalence_100m_parallel_execution():
    """BASELINE configs[1] shape (5M hidden, 100M synapses, 150M-event passes), the throughput kernels themselves
    (PARALLEL: k_traverse_line32 against the iid k_traverse_parallel), 50 passes, 8 seeds."""
    import torch
    if torch.cuda.mem_get_info()[0] < 30 * 2**30:
        pytest.skip("needs 30 GB of free device memory")
    shape = dict(n_input=256, n_output=256, n_hidden=5_000_000, n_syn=100_000_000)
    res = {name: [run_engine(shape, 150_000_000, blk, order, s, capi.EXEC_PARALLEL, passes=50, settle=20) for s in SEEDS]
           for name, (blk, order) in SAMPLERS.items()}
    # Floor 1 % instead of 0.5 %: here each arm also carries the execution-order effect of its own PARALLEL kernel (bounded
    # at 2 % against EXACT execution elsewhere). Measured: line8_interleaved fires +0.55 % against the iid kernel at this
    # shape — 1.2 % of the 12.5M lines of the 100M table are in flight in two warps at once (0.12 % at the 1B shape), and
    # such a pair cannot see each other's fires.
    for line in check_equivalence(res, floor_counts=0.01):
        print(line)
