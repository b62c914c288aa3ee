"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol that
include/abnn.h declares, the ctypes binding covers exactly that set, struct layouts agree between the
header (as compiled by gcc) and the binding, and the device-free entry points behave."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from abnn_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "abnn.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(abnn_[a-z0-9_]+)\s*\(", src))


def test_header_symbols_are_exported_and_bound():
    lib = capi.load()                                    # raises if the library has not been built
    decl = declared_functions()
    assert decl == set(capi.SIGNATURES), (decl ^ set(capi.SIGNATURES))
    exported = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    for name in decl:
        assert re.search(rf"\bT {name}\b", exported), f"{name} is declared in include/abnn.h but not exported"
        getattr(lib, name)


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof/offsetof as gcc sees include/abnn.h == the ctypes mirror."""
    probe = tmp_path / "probe.c"
    fields = {
        "abnn_params": ["struct_size", "n_hidden", "seed", "sampler", "window_pre", "base_scale", "w_prune", "rate_alpha",
                        "filter_tau", "device", "sample_block", "table_order"],
        "abnn_info": ["n_neuron", "neuron_lo", "rank", "l2_bytes", "event_base"],
        "abnn_pass_stats": ["events", "candidates", "clock", "traverse_ms"],
        "abnn_structural_stats": ["n_before", "dropped"],
        "abnn_synapse": ["src", "w"],
    }
    body = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for st, fs in fields.items():
        body.append(f'printf("{st} %zu\\n", sizeof({st}));')
        for f in fs:
            body.append(f'printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
    body.append("return 0;}")
    probe.write_text("\n".join(body))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", str(probe), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    mirror = {"abnn_params": capi.Params, "abnn_info": capi.Info, "abnn_pass_stats": capi.PassStats,
              "abnn_structural_stats": capi.StructuralStats, "abnn_synapse": capi.Synapse}
    for st, fs in fields.items():
        assert int(got[st]) == C.sizeof(mirror[st]), st
        for f in fs:
            assert int(got[f"{st}.{f}"]) == getattr(mirror[st], f).offset, f"{st}.{f}"


def test_default_params_match_the_reference_constants(oracle):
    """abnn_default_params (product) against the independent restatement of constants.h in the oracle."""
    for profile in (capi.PROFILE_METAL_PARITY, capi.PROFILE_NORTH_STAR, capi.PROFILE_B200):
        a, b = capi.default_params(profile), oracle.default_params(profile)
        assert bytes(a) == bytes(b), [n for n, _ in capi.Params._fields_ if n != "reserved_" and getattr(a, n) != getattr(b, n)]


def test_device_free_entry_points(oracle):
    lib = capi.load()
    assert lib.abnn_abi_version() == 1
    lo, hi = C.c_uint64(), C.c_uint64()
    assert lib.abnn_partition(5_000_512, 8, 7, C.byref(lo), C.byref(hi)) == 0
    assert (lo.value, hi.value) == (7 * 625_064, 5_000_512)
    assert lib.abnn_partition(10, 0, 0, C.byref(lo), C.byref(hi)) == capi.ERR_INVALID
    assert b"partition" in lib.abnn_last_error()
    f, c = C.c_uint64(), C.c_uint64()
    total = 0
    for r in range(3):
        assert lib.abnn_event_share(1000, 30, 10 * r, 10, C.byref(f), C.byref(c)) == 0
        total += c.value
    assert total == 1000
    for ctr, key, want in [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
                           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
                            (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]:
        out = (C.c_uint32 * 4)()
        lib.abnn_philox4x32((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert tuple(out) == want == tuple(oracle.philox(ctr, key))


def test_create_without_a_gpu_fails_loudly():
    """No CPU fallback: on a box without a CUDA device abnn_create reports ABNN_ERR_NO_DEVICE."""
    if os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0"):
        pytest.skip("a GPU is present")
    lib = capi.load()
    p = capi.default_params()
    p.n_hidden, p.n_syn = 100, 1000
    h = C.c_void_p()
    assert lib.abnn_create(C.byref(p), C.byref(h)) == capi.ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.abnn_last_error()
    import torch                                         # loading the library first must not break PyTorch's NCCL
    assert not torch.cuda.is_available()


def test_manifest_loader(tmp_path):
    """abnn_params_from_manifest on the reference's flat keys (simple.yml:3-12): underscore numbers,
    comments, nested blocks skipped, reference key names mapped, abnn_params field names accepted."""
    lib = capi.load()
    p = capi.default_params(capi.PROFILE_NORTH_STAR)
    steps, tau_ltd = C.c_uint64(), C.c_uint64()
    path = os.path.join(ROOT, "tests", "golden", "abnn_manifest.yml").encode()
    assert lib.abnn_params_from_manifest(path, C.byref(p), C.byref(steps), C.byref(tau_ltd)) == 0, lib.abnn_last_error()
    assert (p.n_input, p.n_output, p.n_hidden, p.n_syn) == (256, 256, 65536 - 512, 524288)
    assert (p.window_pre, tau_ltd.value, steps.value, p.seed) == (20_000, 40_000, 1_000_000, 42)
    assert (np.float32(p.a_ltp), np.float32(p.a_ltd)) == (np.float32(0.01), np.float32(0.005))
    assert (np.float32(p.w_min), np.float32(p.w_max)) == (np.float32(0.001), np.float32(1.0))      # nested w_max ignored
    assert (p.sample_block, p.table_order) == (8, capi.TABLE_DST_SORTED)                          # abnn_params field names as keys
    bad = tmp_path / "bad.yml"
    bad.write_text("synapses: lots\n")
    assert lib.abnn_params_from_manifest(str(bad).encode(), C.byref(p), None, None) == capi.ERR_INVALID
    assert b"needs a number" in lib.abnn_last_error()
    small = tmp_path / "small.yml"
    small.write_text("neurons: 100\n")
    assert lib.abnn_params_from_manifest(str(small).encode(), C.byref(p), None, None) == capi.ERR_INVALID
    assert lib.abnn_params_from_manifest(b"/nonexistent.yml", C.byref(p), None, None) == capi.ERR_IO


def test_manifest_loader_on_the_reference_manifest():
    """The reference's own abnn/manifests/simple.yml (read where it lies; present in the build container only): the ABNN keys
    at :3-12 map as documented — values the reference's fkYAML-based loader returns as STRINGS (`20_000`) become numbers —
    and the dense-NN blocks below them (layers / training / dataset) are skipped."""
    ref = "/root/reference/abnn/manifests/simple.yml"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present (GPU box)")
    lib = capi.load()
    p = capi.default_params(capi.PROFILE_NORTH_STAR)
    steps, tau_ltd = C.c_uint64(), C.c_uint64()
    assert lib.abnn_params_from_manifest(ref.encode(), C.byref(p), C.byref(steps), C.byref(tau_ltd)) == 0, lib.abnn_last_error()
    assert (p.n_input, p.n_output, p.n_hidden, p.n_syn) == (256, 256, 65536 - 512, 524288)         # neurons: 65536, synapses: 524288
    assert (p.window_pre, tau_ltd.value, steps.value, p.seed) == (20_000, 40_000, 1_000_000, 42)   # tau_LTP / tau_LTD / steps / rng_seed
    assert (np.float32(p.a_ltp), np.float32(p.a_ltd)) == (np.float32(0.01), np.float32(0.005))
    assert (np.float32(p.w_min), np.float32(p.w_max)) == (np.float32(0.001), np.float32(1.0))
    q = capi.default_params(capi.PROFILE_NORTH_STAR)                                                # nothing else moved
    for name, _ in capi.Params._fields_:
        if name not in ("n_hidden", "n_syn", "window_pre", "seed", "a_ltp", "a_ltd", "w_min", "w_max"):
            assert getattr(p, name) == getattr(q, name), name
