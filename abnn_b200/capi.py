"""ctypes binding of include/abnn.h (the C-ABI of libabnn_b200.so).

This is what a host written in another language would bind; the Python
mirror of the reference's `Brain` / `BrainEngine` classes in abnn_b200/brain.py sits on top of it.
The library is the product: loading fails loudly when it has not been built — there is no CPU
or PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ABNN_B200_LIB") or os.path.join(_HERE, "libabnn_b200.so")   # override: tuning variants only

# ---- enums (include/abnn.h) -------------------------------------------------------------------
OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_IO, ERR_SHAPE, ERR_CAPACITY, ERR_COMM, ERR_UNSUPPORTED = (
    0, -1, -2, -3, -4, -5, -6, -7, -8)
SAMPLER_SWEEP, SAMPLER_PHILOX = 0, 1
RNG_XORSHIFT, RNG_PHILOX = 0, 1
CLOCK_PER_PASS, CLOCK_PER_EVENT = 0, 1
EXEC_SERIAL, EXEC_EXACT, EXEC_PARALLEL = 0, 1, 2
SRC_LIVE, SRC_SNAPSHOT = 0, 1
RBAR_PASS_STEP, RBAR_METAL_TID0 = 0, 1
GRAPH_REFERENCE, GRAPH_ER_BETA = 0, 1
TABLE_AS_GIVEN, TABLE_DST_SORTED, TABLE_DST_INTERLEAVED = 0, 1, 2
EXCHANGE_NCCL, EXCHANGE_PEER = 0, 1
PROFILE_METAL_PARITY, PROFILE_NORTH_STAR, PROFILE_B200 = 0, 1, 2


class Synapse(C.Structure):
    """abnn_synapse == SynapsePacked (reference abnn/src/core/brain/brain.h:21)."""
    _fields_ = [("src", C.c_uint32), ("dst", C.c_uint32), ("w", C.c_float), ("pad", C.c_float)]


class Params(C.Structure):
    """abnn_params."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("abi_version", C.c_uint32),
        ("n_input", C.c_uint32), ("n_output", C.c_uint32),
        ("n_hidden", C.c_uint64), ("n_syn", C.c_uint64), ("syn_capacity", C.c_uint64),
        ("seed", C.c_uint64),
        ("sampler", C.c_uint32), ("release_rng", C.c_uint32), ("clock_mode", C.c_uint32),
        ("exec_mode", C.c_uint32), ("src_view", C.c_uint32), ("rbar_mode", C.c_uint32),
        ("max_spikes_per_pass", C.c_uint32), ("track_visits", C.c_uint32),
        ("window_pre", C.c_uint64), ("refractory", C.c_uint64), ("teacher_gap", C.c_uint64),
        ("base_scale", C.c_float), ("a_ltp", C.c_float), ("a_ltd", C.c_float),
        ("w_min", C.c_float), ("w_max", C.c_float), ("eta_home", C.c_float),
        ("target_rate_hz", C.c_float), ("home_tick_hz", C.c_float),
        ("eta_reward", C.c_float), ("alpha_rbar", C.c_float),
        ("w_prune", C.c_float), ("p_new", C.c_float), ("w_init", C.c_float),
        ("rate_alpha", C.c_float), ("peak_decay", C.c_float), ("peak_init", C.c_float),
        ("use_fir", C.c_uint32), ("fir_size", C.c_uint32), ("reward_window", C.c_uint32),
        ("filter_tau", C.c_double), ("dt_sec", C.c_double), ("loss0", C.c_double),
        ("device", C.c_int32), ("rank", C.c_uint32), ("world_size", C.c_uint32),
        ("l2_persist", C.c_uint32),
        ("sample_block", C.c_uint32), ("table_order", C.c_uint32), ("prune_in_place", C.c_uint32), ("exchange", C.c_uint32),
        ("compact_every", C.c_uint32), ("reserved_", C.c_uint32),
    ]

    def copy(self) -> "Params":
        q = Params()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(Params))
        return q


class Info(C.Structure):
    _fields_ = [
        ("n_input", C.c_uint32), ("n_output", C.c_uint32),
        ("n_hidden", C.c_uint64), ("n_neuron", C.c_uint64),
        ("n_syn_global", C.c_uint64), ("n_syn_local", C.c_uint64), ("syn_capacity", C.c_uint64),
        ("neuron_lo", C.c_uint64), ("neuron_hi", C.c_uint64), ("neuron_slice", C.c_uint64),
        ("rank", C.c_uint32), ("world_size", C.c_uint32),
        ("device", C.c_int32), ("sm_count", C.c_uint32),
        ("l2_bytes", C.c_uint64), ("l2_persist_bytes", C.c_uint64),
        ("pass_index", C.c_uint64), ("clock", C.c_uint64), ("event_base", C.c_uint64),
    ]


class PassStats(C.Structure):
    _fields_ = [
        ("events", C.c_uint64), ("gated", C.c_uint64), ("fired", C.c_uint64),
        ("candidates", C.c_uint64), ("grown", C.c_uint64), ("clock", C.c_uint64),
        ("device_ms", C.c_double), ("traverse_ms", C.c_double),
    ]


class StructuralStats(C.Structure):
    _fields_ = [
        ("n_before", C.c_uint64), ("pruned", C.c_uint64), ("appended", C.c_uint64),
        ("n_after", C.c_uint64), ("dropped", C.c_uint64),
    ]


# name -> (restype, argtypes). Every symbol include/abnn.h declares; tests/test_abi.py checks
# this table against the header and against the built library.
_H = C.c_void_p
_P = C.POINTER
SIGNATURES = {
    "abnn_last_error": (C.c_char_p, []),
    "abnn_abi_version": (C.c_uint32, []),
    "abnn_default_params": (C.c_int, [_P(Params), C.c_uint32]),
    "abnn_create": (C.c_int, [_P(Params), _P(_H)]),
    "abnn_destroy": (None, [_H]),
    "abnn_get_info": (C.c_int, [_H, _P(Info)]),
    "abnn_comm_unique_id": (C.c_int, [C.c_void_p]),
    "abnn_comm_init": (C.c_int, [_H, C.c_void_p]),
    "abnn_init_graph": (C.c_int, [_H, C.c_uint32, C.c_uint64]),
    "abnn_upload_synapses": (C.c_int, [_H, C.c_void_p, C.c_uint64]),
    "abnn_download_synapses": (C.c_int, [_H, C.c_void_p, C.c_uint64, _P(C.c_uint64)]),
    "abnn_save_bnn": (C.c_int, [_H, C.c_char_p]),
    "abnn_load_bnn": (C.c_int, [_H, C.c_char_p]),
    "abnn_save_state": (C.c_int, [_H, C.c_char_p]),
    "abnn_load_state": (C.c_int, [_H, C.c_char_p]),
    "abnn_params_from_manifest": (C.c_int, [C.c_char_p, _P(Params), _P(C.c_uint64), _P(C.c_uint64)]),
    "abnn_inject_inputs": (C.c_int, [_H, C.c_void_p, C.c_uint32, C.c_float]),
    "abnn_teacher_force": (C.c_int, [_H, C.c_void_p, C.c_uint32, C.c_float]),
    "abnn_set_reward": (C.c_int, [_H, C.c_float]),
    "abnn_get_reward": (C.c_int, [_H, _P(C.c_float), _P(C.c_float)]),
    "abnn_run_pass": (C.c_int, [_H, C.c_uint64, _P(PassStats)]),
    "abnn_sync": (C.c_int, [_H]),
    "abnn_engine_step": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_uint64, C.c_void_p]),
    "abnn_timer_mark": (C.c_int, [_H, C.c_uint32]),
    "abnn_timer_elapsed": (C.c_int, [_H, C.c_uint32, C.c_uint32, _P(C.c_double)]),
    "abnn_read_outputs": (C.c_int, [_H, C.c_void_p, C.c_uint32]),
    "abnn_readout_filtered": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_uint32]),
    "abnn_readout_step": (C.c_int, [_H, C.c_void_p, C.c_uint32]),
    "abnn_get_loss": (C.c_int, [_H, _P(C.c_double), _P(C.c_uint64)]),
    "abnn_prune_and_grow": (C.c_int, [_H, _P(StructuralStats)]),
    "abnn_download_timestamps": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "abnn_upload_timestamps": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "abnn_download_gate_words": (C.c_int, [_H, C.c_void_p, _P(C.c_uint32)]),
    "abnn_get_clock": (C.c_int, [_H, _P(C.c_uint64)]),
    "abnn_set_clock": (C.c_int, [_H, C.c_uint64]),
    "abnn_partition": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, _P(C.c_uint64), _P(C.c_uint64)]),
    "abnn_event_share": (C.c_int, [C.c_uint64] * 4 + [_P(C.c_uint64), _P(C.c_uint64)]),
    "abnn_philox4x32": (None, [_P(C.c_uint32), _P(C.c_uint32), _P(C.c_uint32)]),
}

_lib = None


class AbnnError(RuntimeError):
    def __init__(self, status: int, where: str, msg: str):
        super().__init__(f"{where} failed with status {status}: {msg}")
        self.status = status


def _preload_nccl() -> None:
    """libabnn_b200.so needs libnccl.so.2 (>= 2.27). A process gets ONE library per soname: if a PyTorch
    wheel with its own, newer NCCL is installed, load that copy first so that a later `import torch`
    (bench.py, abnn_b200.distributed) finds the symbols it was built against. Without such a wheel the
    system libnccl is used. PyTorch itself is never imported here."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia")
    except Exception:
        spec = None
    for root in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(root, "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
            return


def load() -> C.CDLL:
    """Load libabnn_b200.so (built in-tree by __graft_entry__.build()). Raises if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
            "abnn_b200 has no CPU fallback.")
    _preload_nccl()
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.abnn_abi_version() != 1:
        raise ImportError("libabnn_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status: int, where: str) -> None:
    if status != 0:
        msg = load().abnn_last_error()
        raise AbnnError(status, where, msg.decode() if msg else "")


def default_params(profile: int = PROFILE_NORTH_STAR) -> Params:
    p = Params()
    check(load().abnn_default_params(C.byref(p), profile), "abnn_default_params")
    return p


def params_from_manifest(path: str, profile: int = PROFILE_NORTH_STAR):
    """Defaults of `profile` overridden by a flat-key YAML manifest. Returns (params, steps, tau_LTD)."""
    p = default_params(profile)
    steps, tau_ltd = C.c_uint64(), C.c_uint64()
    check(load().abnn_params_from_manifest(path.encode(), C.byref(p), C.byref(steps), C.byref(tau_ltd)), "abnn_params_from_manifest")
    return p, steps.value, tau_ltd.value
