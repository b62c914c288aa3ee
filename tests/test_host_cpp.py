"""The hot path driven from host C++17 through include/abnn_brain.hpp (the façade that mirrors the
reference's Brain / BrainEngine / FunctionalDataset) must reproduce the oracle bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from abnn_b200 import capi
from abnn_b200.build import build_host_tests
from oracle import pyoracle as O


def test_host_cpp_facade_compiles():
    """CPU check: the façade and its driver compile and link against libabnn_b200.so."""
    exe = build_host_tests()
    assert os.path.exists(exe)


def test_p2p_protocol_model(tmp_path):
    """Host model (threads, random delays) of the flag protocol of the opt-in peer-memory exchange (csrc/exchange.cu):
    a rank never reads a gate word of another pass while it traverses, never starts a pass with a stale word, nobody
    deadlocks. Logic only — the CUDA memory-ordering side needs GPUs."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "p2p_protocol_sim"
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-pthread", os.path.join(root, "tests", "host_cpp", "p2p_protocol_sim.cpp"),
                    "-o", str(exe)], check=True, capture_output=True, text=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


@pytest.mark.gpu
def test_host_cpp_engine_matches_oracle(tmp_path):
    exe = build_host_tests()
    out = tmp_path / "engine.bin"
    passes = 30
    r = subprocess.run([exe, str(out), str(passes)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith("ok passes=30")
    raw = np.fromfile(out, dtype=np.uint8)

    over = dict(n_input=64, n_output=64, n_hidden=2000, n_syn=60_000, exec_mode=capi.EXEC_SERIAL,
                clock_mode=capi.CLOCK_PER_PASS, window_pre=5, refractory=2, max_spikes_per_pass=2560, reward_window=10)
    o = O.OracleB(O.default_params(capi.PROFILE_NORTH_STAR, **over))
    o.init_graph(capi.GRAPH_REFERENCE, 1)
    stim = O.Dataset(64, 64)
    even = False
    want = []
    for p in range(passes):
        vin, exp = stim.next_input(), stim.next_expected()
        o.inject_inputs(vin, 1000.0); o.teacher_force(exp, 1.0 if even else 0.0); even = not even
        o.run_pass(60_000)
        want.append(o.readout_filtered(exp).tobytes())
    want.append(o.download_synapses().tobytes())
    assert raw.tobytes() == b"".join(want)
