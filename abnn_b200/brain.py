"""Python mirror of the reference's `Brain` (abnn/src/core/brain/brain.{h,cpp}) and of the per-pass
loop of `BrainEngine` (abnn/src/core/brain-engine.cpp:108-190) on top of the C-ABI.

Same method names and argument meaning as the reference so the parity tests read like tests of the
reference would. Every call goes through libabnn_b200.so (CUDA); nothing here computes on the host
except the stimulus (the reference computes it on the host too: functional-dataset.cpp:24-52).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import capi

SYN_DTYPE = np.dtype([("src", "<u4"), ("dst", "<u4"), ("w", "<f4"), ("pad", "<f4")])

# reference constants (abnn/src/core/constants.h)
INPUT_RATE_HZ = 1000.0
EVENTS_PER_PASS = 150_000_000
INPUT_SIN_WAVE_FREQUENCY = 0.5
DT_SEC = 0.0009


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Brain:
    """Device-state owner. Mirrors Brain's public interface (brain.h:27-58)."""

    def __init__(self, params: capi.Params | None = None, **over):
        self.lib = capi.load()
        p = params.copy() if params is not None else capi.default_params()
        for k, v in over.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        self.params = p
        h = C.c_void_p()
        capi.check(self.lib.abnn_create(C.byref(p), C.byref(h)), "abnn_create")
        self.h = h
        self._n_neuron = p.n_input + p.n_output + p.n_hidden

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.abnn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- getters (brain.h:48-52) ------------------------------------------------------------------
    def info(self) -> capi.Info:
        i = capi.Info()
        capi.check(self.lib.abnn_get_info(self.h, C.byref(i)), "abnn_get_info")
        return i

    def n_input(self): return self.params.n_input
    def n_output(self): return self.params.n_output
    def n_hidden(self): return self.params.n_hidden
    def n_neuron(self): return self._n_neuron
    def n_syn(self): return self.info().n_syn_global

    # -- graph ----------------------------------------------------------------------------------
    def build_random_graph(self, seed: int = 1):
        """build_random_graph (brain-engine.cpp:31-53); the reference seeds mt19937 with 1."""
        capi.check(self.lib.abnn_init_graph(self.h, capi.GRAPH_REFERENCE, seed), "abnn_init_graph")

    def init_graph(self, kind: int, seed: int):
        capi.check(self.lib.abnn_init_graph(self.h, kind, seed), "abnn_init_graph")

    def upload_synapses(self, syn: np.ndarray):
        syn = np.ascontiguousarray(syn, dtype=SYN_DTYPE)
        capi.check(self.lib.abnn_upload_synapses(self.h, _ptr(syn), len(syn)), "abnn_upload_synapses")

    def download_synapses(self) -> np.ndarray:
        n = self.info().n_syn_local
        out = np.zeros(n, SYN_DTYPE)
        got = C.c_uint64()
        capi.check(self.lib.abnn_download_synapses(self.h, _ptr(out), n, C.byref(got)), "abnn_download_synapses")
        return out

    def save(self, path: str):
        capi.check(self.lib.abnn_save_bnn(self.h, path.encode()), "abnn_save_bnn")

    def load(self, path: str):
        capi.check(self.lib.abnn_load_bnn(self.h, path.encode()), "abnn_load_bnn")

    def save_state(self, path: str):
        """.bnn v2: records + timestamps + clock + r-bar + read-out state, for exact resume."""
        capi.check(self.lib.abnn_save_state(self.h, path.encode()), "abnn_save_state")

    def load_state(self, path: str):
        capi.check(self.lib.abnn_load_state(self.h, path.encode()), "abnn_load_state")

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        capi.check(self.lib.abnn_comm_init(self.h, buf), "abnn_comm_init")

    # -- per-pass operations ----------------------------------------------------------------------
    def inject_inputs(self, vals, hz: float = INPUT_RATE_HZ):
        v = np.ascontiguousarray(vals, dtype=np.float32)
        capi.check(self.lib.abnn_inject_inputs(self.h, _ptr(v), len(v), hz), "abnn_inject_inputs")

    def teacher_force(self, expected, rate: float):
        e = np.ascontiguousarray(expected, dtype=np.float32)
        capi.check(self.lib.abnn_teacher_force(self.h, _ptr(e), len(e), rate), "abnn_teacher_force")

    def set_reward(self, r: float):
        capi.check(self.lib.abnn_set_reward(self.h, r), "abnn_set_reward")

    def get_reward(self):
        r, b = C.c_float(), C.c_float()
        capi.check(self.lib.abnn_get_reward(self.h, C.byref(r), C.byref(b)), "abnn_get_reward")
        return np.float32(r.value), np.float32(b.value)

    def encode_traversal(self, events: int):
        """Enqueue one pass (Brain::encode_traversal, brain.cpp:87-122); asynchronous."""
        capi.check(self.lib.abnn_run_pass(self.h, events, None), "abnn_run_pass")

    def run_pass(self, events: int) -> capi.PassStats:
        """One pass + wait (encode_traversal; commit; waitUntilCompleted) with its statistics."""
        st = capi.PassStats()
        capi.check(self.lib.abnn_run_pass(self.h, events, C.byref(st)), "abnn_run_pass")
        return st

    def engine_step(self, vin, expected, hz: float, teacher_rate: float, events: int, want_rates: bool = False):
        """inject + teacher forcing + one pass + read-out step as one enqueue (CUDA graph replay in PARALLEL mode)."""
        v = np.ascontiguousarray(vin, dtype=np.float32)
        e = np.ascontiguousarray(expected, dtype=np.float32)
        out = np.zeros(self.params.n_output, np.float32) if want_rates else None
        capi.check(self.lib.abnn_engine_step(self.h, _ptr(v), _ptr(e), hz, teacher_rate, events, _ptr(out)), "abnn_engine_step")
        return out

    def sync(self):
        capi.check(self.lib.abnn_sync(self.h), "abnn_sync")

    def timer_mark(self, slot: int):
        capi.check(self.lib.abnn_timer_mark(self.h, slot), "abnn_timer_mark")

    def timer_elapsed(self, a: int, b: int) -> float:
        ms = C.c_double()
        capi.check(self.lib.abnn_timer_elapsed(self.h, a, b, C.byref(ms)), "abnn_timer_elapsed")
        return ms.value

    def read_outputs(self) -> np.ndarray:
        out = np.zeros(self.params.n_output, np.uint8)
        capi.check(self.lib.abnn_read_outputs(self.h, _ptr(out), len(out)), "abnn_read_outputs")
        return out

    def readout_filtered(self, expected=None) -> np.ndarray:
        out = np.zeros(self.params.n_output, np.float32)
        e = None if expected is None else np.ascontiguousarray(expected, dtype=np.float32)
        capi.check(self.lib.abnn_readout_filtered(self.h, _ptr(e), _ptr(out), len(out)), "abnn_readout_filtered")
        return out

    def readout_step(self, expected=None):
        e = None if expected is None else np.ascontiguousarray(expected, dtype=np.float32)
        capi.check(self.lib.abnn_readout_step(self.h, _ptr(e), self.params.n_output), "abnn_readout_step")

    def get_loss(self):
        l, w = C.c_double(), C.c_uint64()
        capi.check(self.lib.abnn_get_loss(self.h, C.byref(l), C.byref(w)), "abnn_get_loss")
        return l.value, w.value

    def prune_and_grow(self) -> capi.StructuralStats:
        st = capi.StructuralStats()
        capi.check(self.lib.abnn_prune_and_grow(self.h, C.byref(st)), "abnn_prune_and_grow")
        return st

    # -- raw state (last_fired_buffer()/clock_buffer(), brain.h:54-58) ------------------------------
    def timestamps(self):
        lf = np.zeros(self._n_neuron, np.uint64)
        lv = np.zeros(self._n_neuron, np.uint64)
        capi.check(self.lib.abnn_download_timestamps(self.h, _ptr(lf), _ptr(lv)), "abnn_download_timestamps")
        return lf, lv

    def upload_timestamps(self, lf=None, lv=None):
        lf = None if lf is None else np.ascontiguousarray(lf, dtype=np.uint64)
        lv = None if lv is None else np.ascontiguousarray(lv, dtype=np.uint64)
        capi.check(self.lib.abnn_upload_timestamps(self.h, _ptr(lf), _ptr(lv)), "abnn_upload_timestamps")

    def gate_words(self):
        """(words, valid): the 32-bit pre-spike gate words prepared for the next pass (inspection)."""
        w = np.zeros(self._n_neuron, np.uint32)
        v = C.c_uint32()
        capi.check(self.lib.abnn_download_gate_words(self.h, _ptr(w), C.byref(v)), "abnn_download_gate_words")
        return w, bool(v.value)

    @property
    def clock(self) -> int:
        c = C.c_uint64()
        capi.check(self.lib.abnn_get_clock(self.h, C.byref(c)), "abnn_get_clock")
        return c.value

    @clock.setter
    def clock(self, v: int):
        capi.check(self.lib.abnn_set_clock(self.h, v), "abnn_set_clock")


class FunctionalDataset:
    """StimulusProvider of the reference app (stimulus/functional-dataset.cpp:24-52) with the two
    lambdas installed at view-delegate.cpp:37-42: input cos^2(x), expected 0.5*sin(x)+0.5, where x is
    narrowed to float before the call. Host-side, like the reference."""

    def __init__(self, n_input=256, n_output=256, dt_sec=DT_SEC, freq_hz=INPUT_SIN_WAVE_FREQUENCY):
        self.n_in, self.n_out, self.dt, self.f = n_input, n_output, dt_sec, freq_hz
        self.phase = 0.0
        self.t = 0.0
        self._xi = np.arange(n_input, dtype=np.float64) / n_input
        self._xo = np.arange(n_output, dtype=np.float64) / n_output

    def nextInput(self) -> np.ndarray:
        self.phase += self.f * self.dt
        if self.phase > 1.0:
            self.phase -= 1.0
        self.t += self.dt
        a = (2.0 * math.pi * (self._xi + self.phase)).astype(np.float32).astype(np.float64)
        c = np.cos(a)
        return (c * c).astype(np.float32)

    def nextExpected(self) -> np.ndarray:
        a = (2.0 * math.pi * (self._xo + self.phase)).astype(np.float32).astype(np.float64)
        return (np.float32(0.5) * np.sin(a) + np.float32(0.5)).astype(np.float32)

    def time(self) -> float:
        return self.t


class BrainEngine:
    """The per-pass loop of the reference's BrainEngine::run_one_pass (brain-engine.cpp:108-190):
    stimulus -> inject_inputs -> teacher forcing on alternate passes -> traversal -> read-out filter
    -> windowed loss/reward. The loss/reward step runs on the device inside readout."""

    def __init__(self, brain: Brain, events_per_pass: int = EVENTS_PER_PASS, stimulus=None):
        self.brain = brain
        self.events = events_per_pass
        self.stim = stimulus or FunctionalDataset(brain.n_input(), brain.n_output(), brain.params.dt_sec)
        self._even = False                       # `static bool even = false` (brain-engine.cpp:126)
        self.step = 0

    def set_stimulus(self, stim):
        self.stim = stim

    def run_one_pass(self, want_rates: bool = True):
        b = self.brain
        vin = self.stim.nextInput()
        expected = self.stim.nextExpected()
        rate = 1.0 if self._even else 0.0
        self._even = not self._even
        self.step += 1
        return b.engine_step(vin, expected, INPUT_RATE_HZ, rate, self.events, want_rates)
