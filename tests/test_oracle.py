"""CPU tests that pin the oracles (no GPU, no product code on the path).

Oracle A = the reference's brain.metal compiled verbatim (oracle/oracle_a.cpp); its outputs are
checked against the golden checksums recorded from the reference in SURVEY.md §8c (also stored
in tests/golden/metal_kat.json). Oracle B (oracle/oracle_b.cpp, the restatement every GPU parity
test compares with) is then checked bit-for-bit against Oracle A in the metal-parity profile.
"""
import json
import os

import numpy as np
import pytest

from abnn_b200 import capi

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def kat_graph(n=1024, N=64):
    from oracle.pyoracle import SYN_DTYPE
    i = np.arange(n)
    syn = np.zeros(n, SYN_DTYPE)
    syn["src"] = i % N
    syn["dst"] = (i * 7 + 3) % N
    syn["w"] = (np.float32(0.5) + np.float32(0.4) * ((i % 10).astype(np.float32) / np.float32(10.0))).astype(np.float32)
    return syn


# Philox4x32-10 known-answer vectors (Random123 kat_vectors)
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


@pytest.mark.parametrize("ctr,key,want", PHILOX_KAT)
def test_philox_kat(oracle, ctr, key, want):
    assert tuple(oracle.philox(ctr, key)) == want


def test_oracle_a_matches_survey_goldens(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    gold = json.load(open(os.path.join(GOLD, "metal_kat.json")))
    for variant in ("naive", "hold_clock"):
        a = oracle.OracleA(kat_graph(), 64, reward=0.1, hold_clock=(variant == "hold_clock"))
        for p, g in enumerate(gold[variant]["passes"]):
            a.run_pass(1024)
            assert a.st.budget == g["budget"], (variant, p)
            assert a.st.clock == p + 1
            assert oracle.fnv1a64(a.syn.tobytes()) == g["syn"], (variant, p)
            assert oracle.fnv1a64(a.lastF.tobytes()) == g["lastF"], (variant, p)
            assert np.float32(a.st.rbar) == np.float32(g["rbar"]), (variant, p)
        assert a.lastF[:8].tolist() == gold[variant]["lastF_head"]
        for idx, w in gold[variant]["w_spot"].items():
            assert np.float32(a.syn["w"][int(idx)]) == np.float32(w)


def _metal_params(oracle, n_syn, n_in, n_out, n_hid, **over):
    return oracle.default_params(capi.PROFILE_METAL_PARITY, n_input=n_in, n_output=n_out,
                                 n_hidden=n_hid, n_syn=n_syn, **over)


def _compare_b_to_a(oracle, syn, N, passes, reward, n_in=16, n_out=16, pre=None, events=None):
    a = oracle.OracleA(syn, N, reward=reward, hold_clock=True)
    b = oracle.OracleB(_metal_params(oracle, len(syn), n_in, n_out, N - n_in - n_out))
    b.upload_synapses(syn)
    b.set_reward(reward)
    if pre is not None:
        a.lastF[:] = pre.astype(np.uint32)
        a.st.clock = int(pre.max()) + 1
        b.upload_timestamps(pre.astype(np.uint64))
        b.clock = int(pre.max()) + 1
    ev = events or len(syn)
    for p in range(passes):
        a.run_pass(ev)
        st = b.run_pass(ev)
        sb = b.download_synapses()
        lf, _ = b.timestamps()
        assert sb.tobytes() == a.syn.tobytes(), f"weights differ at pass {p}"
        assert np.array_equal(lf, a.lastF.astype(np.uint64)), f"lastFired differs at pass {p}"
        assert b.clock == a.st.clock
        assert np.float32(b.get_reward()[1]) == np.float32(a.st.rbar)
        assert 2560 - st.fired == a.st.budget, f"budget differs at pass {p}"
    return b


def test_oracle_b_equals_oracle_a_kat(oracle):
    """Metal-parity profile of the restatement == verbatim reference kernel (hold-clock sweep)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    _compare_b_to_a(oracle, kat_graph(), 64, passes=8, reward=0.1)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_b_equals_oracle_a_random(oracle, seed):
    """Random graph, pre-seeded timestamps so all gates, LTP, LTD, reward, budget exhaustion occur."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(seed)
    N, n = 512, 256 * 40
    syn = np.zeros(n, oracle.SYN_DTYPE)
    syn["src"] = rng.integers(0, N, n)
    syn["dst"] = rng.integers(0, N, n)
    syn["w"] = rng.uniform(0.05, 1.0, n).astype(np.float32)
    pre = rng.integers(1, 12, N)
    b = _compare_b_to_a(oracle, syn, N, passes=10, reward=-0.3 + 0.2 * seed, pre=pre)
    assert b.get_reward()[1] != 0  # the tid-0 r-bar quirk was exercised


def test_oracle_b_budget_exhaustion_matches_a(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(7)
    N, n = 4096, 256 * 64
    syn = np.zeros(n, oracle.SYN_DTYPE)
    syn["src"] = rng.integers(0, N, n)
    syn["dst"] = rng.integers(0, N, n)
    syn["w"] = rng.uniform(0.9, 1.0, n).astype(np.float32)   # p ~ 0.7: > 2560 fires wanted
    pre = np.full(N, 1)
    a = oracle.OracleA(syn, N, reward=0.0, hold_clock=True)
    a.lastF[:] = 1
    a.st.clock = 5
    b = oracle.OracleB(_metal_params(oracle, n, 16, 16, N - 32))
    b.upload_synapses(syn)
    b.upload_timestamps(pre.astype(np.uint64))
    b.clock = 5
    a.run_pass(n)
    st = b.run_pass(n)
    assert a.st.budget == 0 and st.fired == 2560
    assert b.download_synapses().tobytes() == a.syn.tobytes()


def test_dataset_and_filter_restatements(oracle):
    """FunctionalDataset / RateFilter restatements == reference sources compiled verbatim == goldens."""
    gold = json.load(open(os.path.join(GOLD, "stimulus_filter.json")))
    d = oracle.Dataset()
    ins, exps = [], []
    for _ in range(gold["frames"]):
        ins.append(d.next_input())
        exps.append(d.next_expected())
    ins, exps = np.array(ins), np.array(exps)
    assert np.array_equal(ins.view(np.uint32), np.array(gold["input_bits"], dtype=np.uint32))
    assert np.array_equal(exps.view(np.uint32), np.array(gold["expected_bits"], dtype=np.uint32))
    # SURVEY.md §8c spot values (printed with %.9g by the survey probe)
    assert "%.9g" % ins[0][0] == "0.999992013" and "%.9g" % ins[0][64] == "7.99445024e-06"
    assert "%.9g" % exps[0][0] == "0.501413703" and "%.9g" % exps[2][0] == "0.504241109"
    if oracle.have_ref():
        R = oracle.RefPieces().L
        h = R.refp_dataset_create(256, 256, 0.0009, 0.5)
        f = R.refp_filter_create(0.02, 1, 20)
        v = np.zeros(256, np.float32)
        o = np.zeros(256, np.float32)
        p = oracle.default_params(n_hidden=16, n_syn=0, use_fir=1)
        for k in range(gold["frames"]):
            R.refp_dataset_next_input(h, v.ctypes.data)
            assert np.array_equal(v, ins[k])
            R.refp_dataset_next_expected(h, v.ctypes.data)
            assert np.array_equal(v, exps[k])


def test_readout_restatement_matches_reference_filter(oracle):
    """ob_readout_filtered's IIR+FIR stage == RateFilter::process (verbatim) on the same spike trains."""
    gold = json.load(open(os.path.join(GOLD, "stimulus_filter.json")))
    p = oracle.default_params(n_hidden=64, n_syn=0, clock_mode=capi.CLOCK_PER_PASS, src_view=capi.SRC_LIVE)
    b = oracle.OracleB(p)
    rng = np.random.default_rng(5)
    N = 256 + 256 + 64
    have = oracle.have_ref()
    if have:
        R = oracle.RefPieces().L
        f = R.refp_filter_create(0.02, 1, 20)
    rate = np.zeros(256, np.float32)
    maxobs = np.float32(0.5)
    outs = []
    for k in range(40):
        spikes = rng.random(256) < (0.2 + 0.6 * (k % 7) / 7)
        b.clock = k + 2
        lf = np.zeros(N, np.uint64)
        lf[256:512] = np.where(spikes, k + 1, 0)    # ts in [now-1, now) <=> spike (ts == 0 never counts)
        b.upload_timestamps(lf)
        got = b.readout_filtered(None)
        outs.append(got)
        if have:
            sp = b.read_outputs().astype(bool)
            rate = (np.float32(0.5) * rate + np.float32(0.5) * sp.astype(np.float32)).astype(np.float32)
            sm = np.zeros(256, np.float32)
            R.refp_filter_process(f, rate.ctypes.data, 256, 0.0009, sm.ctypes.data)
            maxobs = np.float32(max(maxobs, sm.max()) * np.float32(0.999))
            want = np.minimum(sm / maxobs, np.float32(1.0)).astype(np.float32)
            assert np.array_equal(got, want), k
    assert np.array_equal(np.array(outs).view(np.uint32), np.array(gold["readout_bits"], dtype=np.uint32))


def test_dst_sorted_table_is_a_stable_sort(oracle):
    """ABNN_TABLE_DST_SORTED: the oracle's table equals numpy's stable sort by dst after upload and
    after a growth step (new records land behind the existing ones of their destination)."""
    rng = np.random.default_rng(4)
    n, N = 50_000, 700
    syn = np.zeros(n, oracle.SYN_DTYPE)
    syn["src"] = rng.integers(0, N, n); syn["dst"] = rng.integers(0, N, n)
    syn["w"] = rng.uniform(0.3, 1.0, n).astype(np.float32)
    p = oracle.default_params(capi.PROFILE_NORTH_STAR, n_input=8, n_output=8, n_hidden=N - 16, n_syn=n,
                              exec_mode=capi.EXEC_SERIAL, table_order=capi.TABLE_DST_SORTED, sample_block=8,
                              window_pre=10**9, refractory=5, p_new=0.5, syn_capacity=n + 10_000)
    o = oracle.OracleB(p)
    o.upload_synapses(syn)
    want = syn[np.argsort(syn["dst"], kind="stable")]
    assert o.download_synapses().tobytes() == want.tobytes()
    o.upload_timestamps(np.full(N, 1, np.uint64), None); o.clock = 100
    st = o.run_pass(40_000)
    assert st.fired > 100 and st.grown > 50
    before = o.download_synapses()
    ss = o.prune_and_grow()
    after = o.download_synapses()
    assert ss.appended == st.grown and len(after) == n + ss.appended
    assert np.all(np.diff(after["dst"].astype(np.int64)) >= 0)
    # existing records keep their relative order; the new ones (w == w_init) sit at the end of their destination's run
    old = after[after["w"] != np.float32(0.1)]
    assert old.tobytes() == before[before["w"] != np.float32(0.1)].tobytes()


def test_oracle_b_matches_frozen_northstar_checksums(oracle):
    """The north-star semantics have no reference implementation to pin against ("parity unpinned"):
    Oracle B is their definition. tests/golden/northstar_oracle.json freezes that definition — BASELINE configs[0]
    (reference graph, iid sampler, sine input, 1M-event passes), the round-1 throughput configuration (line sampler,
    dst-sorted table, pruning + growth every pass), the same on two dst-shards, and the layout bench.py runs
    (ABNN_PROFILE_B200: 16-record groups over the interleaved table) with compact_every = 3 — so a change of oracle_b.cpp,
    of its compile flags or of the host toolchain that alters any result is caught here, on the CPU."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_northstar_golden", os.path.join(GOLD, "make_northstar_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    want = json.load(open(os.path.join(GOLD, "northstar_oracle.json")))
    assert set(want) == set(gen.CASES)
    for name, fn in gen.CASES.items():
        got = json.loads(json.dumps(fn()))
        assert got == want[name], f"{name}: Oracle B no longer reproduces its frozen checksums"
    # the cases are not degenerate: events gate, fire, grow and get pruned
    last = want["toy_line_sorted"][-1]
    assert last["stats"]["gated"] > 100_000 and last["stats"]["fired"] > 1000 and last["structural"]["pruned"] > 0
    assert want["toy_reference"][-1]["structural"]["appended"] > 0
    lazy = want["toy_b200_lazy"]
    assert all(r["structural"]["appended"] > 0 for r in lazy) and lazy[1]["structural"]["pruned"] > 0 and lazy[3]["structural"]["pruned"] > 0


def test_logger_matches_the_reference_logger_byte_for_byte(tmp_path):
    """abnn_b200::Logger (include/abnn_brain.hpp) against the reference's Logger compiled verbatim (logger.cpp:13-84 through
    oracle/ref_pieces.cpp): the same script of 3 frames, 10 losses (the 10th truncates the file, logger.cpp:68), 1 more
    frame; abnn_session.m must be byte-identical after EVERY call and the loss EMA must agree."""
    import struct
    import subprocess
    from oracle import pyoracle as O
    if not O.have_ref():
        pytest.skip("oracle/_ref not built (needs the reference tree)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "logger_parity"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "host_cpp", "logger_parity.cpp"),
                    "-o", str(exe), "-L", os.path.join(root, "abnn_b200"), "-labnn_b200", "-Wl,-rpath," + os.path.join(root, "abnn_b200"), "-pthread"],
                   check=True, capture_output=True, text=True)
    n_in, n_out = 12, 9
    rng = np.random.default_rng(3)
    ops = [("frame", rng.random(n_in).astype(np.float32) * 2 - 1, rng.random(n_out).astype(np.float32)) for _ in range(3)]
    ops += [("loss", 0.25 / (l + 1)) for l in range(10)]
    ops += [("frame", np.linspace(-1, 1, n_in).astype(np.float32), np.full(n_out, 1e-7, np.float32))]
    script = tmp_path / "script.bin"
    with open(script, "wb") as f:
        for op in ops:
            if op[0] == "frame":
                f.write(struct.pack("<i", 0)); f.write(op[1].tobytes()); f.write(op[2].tobytes())
            else:
                f.write(struct.pack("<id", 1, op[1]))
    ours = tmp_path / "ours.m"
    r = subprocess.run([str(exe), str(ours), str(script), str(n_in), str(n_out)], capture_output=True, text=True, check=True)
    ref_exe = tmp_path / "logger_ref_driver"
    ref_lib_dir = os.path.dirname(O.LIB_P)
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(root, "tests", "host_cpp", "logger_ref_driver.cpp"), "-o", str(ref_exe),
                    O.LIB_P, "-Wl,-rpath," + ref_lib_dir], check=True, capture_output=True, text=True)
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    # the reference writes into its current directory (logger.cpp:14): a process of its own, started there
    subprocess.run([str(ref_exe), str(script), str(n_in), str(n_out)], cwd=ref_dir, capture_output=True, text=True, check=True)
    for k, op in enumerate(ops):
        want = (ref_dir / f"ref.{k}").read_bytes()
        got = (tmp_path / f"ours.m.{k}").read_bytes()
        assert got == want, f"abnn_session.m differs after call {k} ({op[0]}): {len(got)} vs {len(want)} bytes"
        if k == 2:
            assert want.count(b"clf;") == 3
        if k == 12:
            assert want == b""                          # truncated by the 10th loss (logger.cpp:68); the new header is still buffered
        if k == 13:
            assert want.startswith(b"% ABNN animated session\n") and want.count(b"clf;") == 1
    ema = 0.0
    for l in range(10):
        ema = 0.25 / (l + 1) if l == 0 else 0.98 * ema + (1.0 - 0.98) * 0.25 / (l + 1)          # logger.cpp:61-62
    assert abs(float(r.stdout.strip().splitlines()[-1].split()[1]) - ema) < 1e-15


def test_oracle_periodic_rebuild_schedule():
    """compact_every = 3 in the oracle: structural steps 0, 3, 6 rebuild (no dead record, table in dst order), the steps in
    between keep every slot (dead records stay, grown synapses sit behind the table) and only the rebuild shrinks it."""
    from oracle import pyoracle as O
    rng = np.random.default_rng(3)
    N, n = 2000, 40_000
    syn = np.zeros(n, O.SYN_DTYPE)
    syn["src"] = rng.integers(0, N, n); syn["dst"] = rng.integers(16, N, n); syn["w"] = rng.uniform(0.15, 1.0, n).astype(np.float32)
    p = O.default_params(capi.PROFILE_NORTH_STAR, n_input=8, n_output=8, n_hidden=N - 16, n_syn=n, exec_mode=capi.EXEC_SERIAL, sample_block=8,
                         table_order=capi.TABLE_DST_SORTED, window_pre=400_000, refractory=5_000, p_new=0.3, w_prune=0.16, w_init=0.17,
                         syn_capacity=n + 20_000, compact_every=3)
    o = O.OracleB(p)
    o.upload_synapses(syn); o.upload_timestamps(rng.integers(1, 20_000, N).astype(np.uint64), None); o.clock = 20_000; o.set_reward(0.1)
    dead_total = 0
    for step in range(7):
        o.run_pass(30_000)
        before = o.download_synapses()
        st = o.prune_and_grow()
        t = o.download_synapses()
        dead = int((t["src"] == 0xFFFFFFFF).sum())
        if step % 3 == 0:
            assert dead == 0 and np.all(np.diff(t["dst"].astype(np.int64)) >= 0)
            assert st.n_after == len(t) <= st.n_before + st.appended
        else:
            assert st.n_after == st.n_before + st.appended == len(t)
            assert t[:len(before)]["dst"].tobytes() == before["dst"].tobytes()       # nothing moved
            assert np.all(t[len(before):]["w"] == np.float32(0.17))                  # the tail holds the grown synapses
            dead_total += dead
    assert dead_total > 0
