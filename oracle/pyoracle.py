"""oracle/pyoracle.py — ctypes front-end of the CPU oracles. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
  * OracleB  : oracle/_build/liboracle_b.so  — host restatement of the traversal (oracle_b.cpp)
  * OracleA  : oracle/_ref/liboracle_a.so    — the reference's brain.metal compiled verbatim
  * RefPieces: oracle/_ref/libref_pieces.so  — the reference's RateFilter / FunctionalDataset
The struct layouts are shared with the product binding (abnn_b200.capi: plain type definitions);
the default parameter values are restated here from the reference's constants on purpose, so the
product's abnn_default_params is checked against an independent copy.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from abnn_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_B = os.path.join(HERE, "_build", "liboracle_b.so")
LIB_A = os.path.join(HERE, "_ref", "liboracle_a.so")
LIB_P = os.path.join(HERE, "_ref", "libref_pieces.so")

SYN_DTYPE = np.dtype([("src", "<u4"), ("dst", "<u4"), ("w", "<f4"), ("pad", "<f4")])
GROW_DTYPE = np.dtype([("order", "<u8"), ("src", "<u4"), ("dst", "<u4")])


def build(force: bool = False) -> None:
    """Compile the oracles (oracle/Makefile). _ref targets are built only where /root/reference exists."""
    if force or not os.path.exists(LIB_B) or os.path.getmtime(LIB_B) < os.path.getmtime(os.path.join(HERE, "oracle_b.cpp")):
        subprocess.check_call(["make", "-C", HERE, "-s", "all"])
    elif os.path.isdir("/root/reference") and not (os.path.exists(LIB_A) and os.path.exists(LIB_P)):
        subprocess.check_call(["make", "-C", HERE, "-s", "ref"])


def default_params(profile: int = capi.PROFILE_NORTH_STAR, **over) -> capi.Params:
    """Reference constants (constants.h:2-19, brain.metal:22-31, brain.h:17-19, brain-engine.h:54,81-84)."""
    p = capi.Params()
    p.struct_size = C.sizeof(capi.Params)
    p.abi_version = 1
    p.n_input, p.n_output, p.n_hidden, p.n_syn = 256, 256, 5_000_000, 1_000_000_000
    p.syn_capacity = 0
    p.seed = 42                                     # manifests/simple.yml:12
    if profile == capi.PROFILE_METAL_PARITY:
        p.sampler, p.release_rng = capi.SAMPLER_SWEEP, capi.RNG_XORSHIFT
        p.clock_mode, p.exec_mode = capi.CLOCK_PER_PASS, capi.EXEC_SERIAL
        p.src_view, p.rbar_mode = capi.SRC_LIVE, capi.RBAR_METAL_TID0
        p.max_spikes_per_pass, p.track_visits = 2560, 0
        p.window_pre, p.refractory = 5, 2
    else:
        p.sampler, p.release_rng = capi.SAMPLER_PHILOX, capi.RNG_PHILOX
        p.clock_mode, p.exec_mode = capi.CLOCK_PER_EVENT, capi.EXEC_PARALLEL
        p.src_view, p.rbar_mode = capi.SRC_SNAPSHOT, capi.RBAR_PASS_STEP
        p.max_spikes_per_pass, p.track_visits = 0, 1
        p.window_pre, p.refractory = 50_000, 2      # brain.cpp:102 tauPre; brain.metal:23
    p.teacher_gap = 1
    p.base_scale, p.a_ltp, p.a_ltd, p.w_min, p.w_max = 0.8, 0.04, 0.02, 0.001, 1.0
    p.eta_home, p.target_rate_hz, p.home_tick_hz = 1.0e-6, 1000.0, 1.0e6
    p.eta_reward, p.alpha_rbar = 1.0e-3, 0.001
    p.w_prune, p.p_new, p.w_init = 0.0, 0.0, 0.1
    p.rate_alpha, p.peak_decay, p.peak_init = 0.5, 0.999, 0.5
    p.use_fir, p.fir_size, p.reward_window = 1, 20, 1000
    p.filter_tau, p.dt_sec, p.loss0 = 0.02, 0.0009, 0.25
    p.device, p.rank, p.world_size, p.l2_persist = -1, 0, 1, 1
    p.sample_block = 1
    if profile == capi.PROFILE_B200:
        p.sample_block, p.table_order = 16, capi.TABLE_DST_INTERLEAVED
    for k, v in over.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


_libb = None


def libb() -> C.CDLL:
    global _libb
    if _libb is None:
        build()
        L = C.CDLL(LIB_B)
        H = C.c_void_p
        L.ob_create.restype = H
        L.ob_create.argtypes = [C.POINTER(capi.Params)]
        L.ob_destroy.argtypes = [H]
        L.ob_live_ptr.restype = C.POINTER(C.c_uint64)
        L.ob_view_ptr.restype = C.POINTER(C.c_uint64)
        L.ob_live_ptr.argtypes = [H]
        L.ob_view_ptr.argtypes = [H]
        for n in ("ob_n_syn_local", "ob_grow_count", "ob_prune"):
            getattr(L, n).restype = C.c_uint64
            getattr(L, n).argtypes = [H]
        L.ob_grow_apply.restype = C.c_uint64
        L.ob_grow_apply.argtypes = [H, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.ob_grow_fetch.argtypes = [H, C.c_void_p]
        L.ob_set_shard_counts.argtypes = [H, C.c_void_p]
        L.ob_upload_synapses.argtypes = [H, C.c_void_p, C.c_uint64]
        L.ob_download_synapses.argtypes = [H, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.ob_init_graph.argtypes = [H, C.c_uint32, C.c_uint64]
        L.ob_inject_inputs.argtypes = [H, C.c_void_p, C.c_uint32, C.c_float]
        L.ob_teacher_force.argtypes = [H, C.c_void_p, C.c_uint32, C.c_float]
        L.ob_set_reward.argtypes = [H, C.c_float]
        L.ob_get_reward.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ob_run_pass.argtypes = [H, C.c_uint64, C.POINTER(capi.PassStats)]
        L.ob_read_outputs.argtypes = [H, C.c_void_p, C.c_uint32]
        L.ob_readout_filtered.argtypes = [H, C.c_void_p, C.c_void_p, C.c_uint32]
        L.ob_get_loss.argtypes = [H, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.ob_prune_and_grow.argtypes = [H, C.POINTER(capi.StructuralStats)]
        L.ob_download_timestamps.argtypes = [H, C.c_void_p, C.c_void_p]
        L.ob_upload_timestamps.argtypes = [H, C.c_void_p, C.c_void_p]
        L.ob_get_clock.argtypes = [H, C.POINTER(C.c_uint64)]
        L.ob_set_clock.argtypes = [H, C.c_uint64]
        L.ob_world_run_pass.argtypes = [C.POINTER(H), C.c_uint32, C.c_uint64, C.POINTER(capi.PassStats)]
        L.ob_world_prune_and_grow.argtypes = [C.POINTER(H), C.c_uint32, C.POINTER(capi.StructuralStats)]
        L.ob_partition.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.ob_event_share.argtypes = [C.c_uint64] * 4 + [C.POINTER(C.c_uint64)] * 2
        L.ob_philox4x32.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.ob_dataset_create.restype = H
        L.ob_dataset_create.argtypes = [C.c_uint32, C.c_uint32, C.c_double, C.c_double]
        L.ob_dataset_destroy.argtypes = [H]
        L.ob_dataset_next_input.argtypes = [H, C.c_void_p]
        L.ob_dataset_next_expected.argtypes = [H, C.c_void_p]
        L.ob_dataset_time.restype = C.c_double
        L.ob_dataset_time.argtypes = [H]
        _libb = L
    return _libb


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    libb().ob_philox4x32(c, k, o)
    return list(o)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleB:
    """One shard of the host restatement; method names follow the C-ABI (include/abnn.h)."""

    def __init__(self, params: capi.Params):
        self.L = libb()
        self.p = params.copy()
        self.h = self.L.ob_create(C.byref(self.p))
        if not self.h:
            raise ValueError("ob_create rejected the parameters")
        self.N = self.p.n_input + self.p.n_output + self.p.n_hidden

    def close(self):
        if self.h:
            self.L.ob_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, where):
        if rc != 0:
            raise RuntimeError(f"oracle {where} -> {rc}")

    def init_graph(self, kind, seed):
        self._ck(self.L.ob_init_graph(self.h, kind, seed), "init_graph")

    def upload_synapses(self, syn: np.ndarray):
        syn = np.ascontiguousarray(syn, dtype=SYN_DTYPE)
        self._ck(self.L.ob_upload_synapses(self.h, _ptr(syn), len(syn)), "upload_synapses")

    def n_syn_local(self) -> int:
        return int(self.L.ob_n_syn_local(self.h))

    def download_synapses(self) -> np.ndarray:
        n = self.n_syn_local()
        out = np.zeros(n, SYN_DTYPE)
        got = C.c_uint64()
        self._ck(self.L.ob_download_synapses(self.h, _ptr(out), n, C.byref(got)), "download_synapses")
        return out

    def set_shard_counts(self, counts):
        a = np.ascontiguousarray(counts, dtype=np.uint64)
        self._ck(self.L.ob_set_shard_counts(self.h, _ptr(a)), "set_shard_counts")

    def inject_inputs(self, v, hz):
        v = np.ascontiguousarray(v, dtype=np.float32)
        self._ck(self.L.ob_inject_inputs(self.h, _ptr(v), len(v), hz), "inject_inputs")

    def teacher_force(self, expected, rate):
        e = np.ascontiguousarray(expected, dtype=np.float32)
        self._ck(self.L.ob_teacher_force(self.h, _ptr(e), len(e), rate), "teacher_force")

    def set_reward(self, r):
        self.L.ob_set_reward(self.h, r)

    def get_reward(self):
        r, b = C.c_float(), C.c_float()
        self.L.ob_get_reward(self.h, C.byref(r), C.byref(b))
        return np.float32(r.value), np.float32(b.value)

    def run_pass(self, events) -> capi.PassStats:
        st = capi.PassStats()
        self._ck(self.L.ob_run_pass(self.h, events, C.byref(st)), "run_pass")
        return st

    def read_outputs(self) -> np.ndarray:
        out = np.zeros(self.p.n_output, np.uint8)
        self._ck(self.L.ob_read_outputs(self.h, _ptr(out), len(out)), "read_outputs")
        return out

    def readout_filtered(self, expected=None) -> np.ndarray:
        out = np.zeros(self.p.n_output, np.float32)
        e = None if expected is None else np.ascontiguousarray(expected, dtype=np.float32)
        self._ck(self.L.ob_readout_filtered(self.h, _ptr(e), _ptr(out), len(out)), "readout_filtered")
        return out

    def get_loss(self):
        l, w = C.c_double(), C.c_uint64()
        self.L.ob_get_loss(self.h, C.byref(l), C.byref(w))
        return l.value, w.value

    def prune_and_grow(self) -> capi.StructuralStats:
        st = capi.StructuralStats()
        self._ck(self.L.ob_prune_and_grow(self.h, C.byref(st)), "prune_and_grow")
        return st

    def prune(self) -> int:
        return int(self.L.ob_prune(self.h))

    def grow_fetch(self) -> np.ndarray:
        n = int(self.L.ob_grow_count(self.h))
        out = np.zeros(n, GROW_DTYPE)
        if n:
            self.L.ob_grow_fetch(self.h, _ptr(out))
        return out

    def grow_apply(self, cands: np.ndarray):
        c = np.ascontiguousarray(cands, dtype=GROW_DTYPE)
        d = C.c_uint64()
        app = self.L.ob_grow_apply(self.h, _ptr(c), len(c), C.byref(d))
        return int(app), int(d.value)

    def timestamps(self):
        lf = np.zeros(self.N, np.uint64)
        lv = np.zeros(self.N, np.uint64)
        self.L.ob_download_timestamps(self.h, _ptr(lf), _ptr(lv))
        return lf, lv

    def upload_timestamps(self, lf=None, lv=None):
        lf = None if lf is None else np.ascontiguousarray(lf, dtype=np.uint64)
        lv = None if lv is None else np.ascontiguousarray(lv, dtype=np.uint64)
        self.L.ob_upload_timestamps(self.h, _ptr(lf), _ptr(lv))

    def live_view(self):
        """numpy views of the shard's live / snapshot lastFired arrays (for the exchange step)."""
        live = np.ctypeslib.as_array(self.L.ob_live_ptr(self.h), shape=(self.N,))
        view = np.ctypeslib.as_array(self.L.ob_view_ptr(self.h), shape=(self.N,))
        return live, view

    @property
    def clock(self) -> int:
        c = C.c_uint64()
        self.L.ob_get_clock(self.h, C.byref(c))
        return c.value

    @clock.setter
    def clock(self, v):
        self.L.ob_set_clock(self.h, v)


class OracleWorld:
    """G shards in one process, one thread per shard per pass (the threaded CPU baseline)."""

    def __init__(self, params: capi.Params, world: int):
        self.shards = []
        for k in range(world):
            q = params.copy()
            q.rank, q.world_size = k, world
            self.shards.append(OracleB(q))
        self.G = world
        self._arr = (C.c_void_p * world)(*[s.h for s in self.shards])
        self.L = libb()

    def init_graph(self, kind, seed):
        for s in self.shards:
            s.init_graph(kind, seed)
        self._sync_counts()

    def upload_synapses(self, syn):
        for s in self.shards:
            s.upload_synapses(syn)
        self._sync_counts()

    def _sync_counts(self):
        counts = [s.n_syn_local() for s in self.shards]
        for s in self.shards:
            s.set_shard_counts(counts)

    def run_pass(self, events) -> capi.PassStats:
        st = capi.PassStats()
        self.L.ob_world_run_pass(self._arr, self.G, events, C.byref(st))
        return st

    def prune_and_grow(self) -> capi.StructuralStats:
        st = capi.StructuralStats()
        self.L.ob_world_prune_and_grow(self._arr, self.G, C.byref(st))
        return st

    def each(self, fn, *a):
        return [getattr(s, fn)(*a) for s in self.shards]


class Dataset:
    """Oracle restatement of FunctionalDataset + the app's lambdas."""

    def __init__(self, n_in=256, n_out=256, dt=0.0009, f=0.5):
        self.L = libb()
        self.n_in, self.n_out = n_in, n_out
        self.h = self.L.ob_dataset_create(n_in, n_out, dt, f)

    def next_input(self):
        v = np.zeros(self.n_in, np.float32)
        self.L.ob_dataset_next_input(self.h, _ptr(v))
        return v

    def next_expected(self):
        v = np.zeros(self.n_out, np.float32)
        self.L.ob_dataset_next_expected(self.h, _ptr(v))
        return v

    def __del__(self):
        try:
            self.L.ob_dataset_destroy(self.h)
        except Exception:
            pass


# ---- Oracle A / reference pieces (only where oracle/_ref was built) ------------------------------
def have_ref() -> bool:
    build()
    return os.path.exists(LIB_A) and os.path.exists(LIB_P)


class _AState(C.Structure):
    _fields_ = [("syn", C.c_void_p), ("lastF", C.c_void_p), ("lastV", C.c_void_p),
                ("clock", C.c_uint32), ("budget", C.c_uint32), ("reward", C.c_float), ("rbar", C.c_float)]


class OracleA:
    """The reference's monte_carlo_traversal (brain.metal:41-130) compiled verbatim, swept serially."""

    def __init__(self, syn: np.ndarray, n_neuron: int, reward=0.0, hold_clock=True):
        self.L = C.CDLL(LIB_A)
        self.L.oracle_a_pass.argtypes = [C.POINTER(_AState)] + [C.c_uint32] * 3 + [C.c_float] * 4 + [C.c_int]
        self.L.oracle_a_renorm.argtypes = [C.POINTER(_AState), C.c_uint32]
        self.syn = np.ascontiguousarray(syn, dtype=SYN_DTYPE).copy()
        self.lastF = np.zeros(n_neuron, np.uint32)
        self.lastV = np.zeros(n_neuron, np.uint32)
        self.st = _AState(self.syn.ctypes.data, self.lastF.ctypes.data, self.lastV.ctypes.data, 0, 0, reward, 0.0)
        self.hold = 1 if hold_clock else 0

    def run_pass(self, events, max_spikes=2560, a_ltp=0.04, a_ltd=0.02, w_min=0.001, w_max=1.0):
        self.L.oracle_a_pass(C.byref(self.st), len(self.syn), events, max_spikes, a_ltp, a_ltd, w_min, w_max, self.hold)

    def renorm(self):
        self.L.oracle_a_renorm(C.byref(self.st), len(self.lastF))


class RefPieces:
    def __init__(self):
        L = C.CDLL(LIB_P)
        H = C.c_void_p
        L.refp_dataset_create.restype = H
        L.refp_dataset_create.argtypes = [C.c_uint32, C.c_uint32, C.c_double, C.c_double]
        L.refp_dataset_next_input.argtypes = [H, C.c_void_p]
        L.refp_dataset_next_expected.argtypes = [H, C.c_void_p]
        L.refp_dataset_destroy.argtypes = [H]
        L.refp_filter_create.restype = H
        L.refp_filter_create.argtypes = [C.c_double, C.c_int, C.c_uint64]
        L.refp_filter_process.argtypes = [H, C.c_void_p, C.c_uint32, C.c_double, C.c_void_p]
        L.refp_filter_destroy.argtypes = [H]
        self.L = L


def fnv1a64(b: bytes) -> str:
    """FNV-1a 64 over raw bytes (the checksum SURVEY.md §8c records its golden values in)."""
    a = np.frombuffer(b, dtype=np.uint8)
    h = 0xCBF29CE484222325
    for x in a.tolist():
        h = ((h ^ x) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h
