"""Shared helpers of the parity tests: seeded inputs and side-by-side drivers for the CUDA path
(abnn_b200.Brain -> C-ABI) and the oracle (oracle.pyoracle.OracleB)."""
import numpy as np

from abnn_b200 import capi
from oracle import pyoracle as O


def random_graph(rng, n_syn, n_neuron, wlo=0.05, whi=1.0, dst_lo=0):
    syn = np.zeros(n_syn, O.SYN_DTYPE)
    syn["src"] = rng.integers(0, n_neuron, n_syn)
    syn["dst"] = rng.integers(dst_lo, n_neuron, n_syn)
    syn["w"] = rng.uniform(wlo, whi, n_syn).astype(np.float32)
    return syn


def assert_same_state(b, o, what=""):
    """Bit-exact comparison of the CUDA handle `b` and the oracle `o`."""
    sb, so = b.download_synapses(), o.download_synapses()
    assert len(sb) == len(so), f"{what}: table length {len(sb)} vs {len(so)}"
    if sb.tobytes() != so.tobytes():
        bad = np.flatnonzero((sb["w"].view(np.uint32) != so["w"].view(np.uint32)) | (sb["src"] != so["src"]) | (sb["dst"] != so["dst"]))
        raise AssertionError(f"{what}: {len(bad)} synapse records differ, first {bad[:5]}: gpu {sb[bad[:3]]} oracle {so[bad[:3]]}")
    lfb, lvb = b.timestamps()
    lfo, lvo = o.timestamps()
    assert np.array_equal(lfb, lfo), f"{what}: lastFired differs at {np.flatnonzero(lfb != lfo)[:8]}"
    assert np.array_equal(lvb, lvo), f"{what}: lastVisited differs at {np.flatnonzero(lvb != lvo)[:8]}"
    assert b.clock == o.clock, f"{what}: clock {b.clock} vs {o.clock}"
    rb, ro = b.get_reward(), o.get_reward()
    assert rb[0].tobytes() == ro[0].tobytes() and rb[1].tobytes() == ro[1].tobytes(), f"{what}: reward/rbar {rb} vs {ro}"


def assert_same_stats(sb, so, what=""):
    for f in ("events", "gated", "fired", "candidates", "grown", "clock"):
        assert getattr(sb, f) == getattr(so, f), f"{what}: stats.{f} {getattr(sb, f)} vs {getattr(so, f)}"
