#!/usr/bin/env python
"""bench.py — synaptic events/s of the ABNN traversal hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--structural]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one engine pass over one synthetic stimulus frame: inject sine input -> teacher forcing -> EVENTS_PER_PASS
traversal events -> (N>1) exchange of the fired-neuron timestamps -> FIR read-out and reward step. Workload =
BASELINE.json configs[2] (the shape the metric is quoted on): 5,000,000 hidden + 256 in + 256 out neurons, 1,000,000,000
synapses (16 GB SynapsePacked), 150,000,000 events per pass; for N>1 the same table is dst-sharded over the ranks
(configs[3], strong scaling). --structural runs BASELINE configs[4] instead: pruning + synaptogenesis after EVERY pass,
4,000,000,000 synapses, 5M hidden neurons and 150M-event passes at N=8; at smaller N the same per-GPU load (500M synapses,
625k hidden neurons and 18.75M events per pass and GPU), weak scaling.

  value    : whole-job events/s, device-timed (CUDA events on the handle's stream), state resident in HBM.
  e2e      : the same through the reference-facing per-pass API with HOST buffers: stimulus vectors are copied
             host->device and the filtered read-out device->host every pass, wall clock between syncs.
  roofline : traversal kernel alone: events * B_alg / kernel time vs MEASURED_PEAKS.json hbm_gbs, B_alg = 16 B + 16 B *
             gated fraction; traffic = ncu DRAM bytes of that kernel (profiles/r2_traffic.json, valid only
             for the kernel source it was captured from), dram_frac = traffic / kernel time / peak.
  samplers / exact_mode : the same workload under the iid sampler of SURVEY §8.0, the dst-sorted line sampler, the
             256-byte block sampler and in EXACT (conflict-free, bit-exact) execution — sub-records beside the headline.
  parity   : (N>1) checked after the timed region on fresh handles: identical gate words on every rank, equal to
             slack_word(lastFired); lastVisited and the first pass's candidates equal to an EXACT-mode run of the same passes.
  cpu_baseline : the oracle (host C++ restatement, oracle/oracle_b.cpp) on all host threads, bounded sample.
--impl reference : the reference ships no CPU traversal and its Metal/AppKit app cannot be built here; this
  arm times the oracle port of the reference algorithm on the host cores: K steps, each a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_IN, N_OUT = 256, 256
HBM_FALLBACK_GBS = 6650.0
TABLE_ORDERS = ("interleaved", "dst", "given")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--hidden", type=int, default=0, help="hidden neurons (default: 5M; --structural: 5M * gpus / 8, the per-GPU load of configs[4])")
    ap.add_argument("--syn", type=int, default=0, help="global synapse count (default: 1e9; --structural: 5e8 per GPU)")
    ap.add_argument("--events", type=int, default=0, help="events per pass (default: 150M; --structural: 150M * gpus / 8, the per-GPU load of configs[4])")
    ap.add_argument("--sampler", default="philox", choices=["philox", "sweep"])
    ap.add_argument("--block", type=int, default=16, help="PHILOX sampler granularity in records (16 = 256 bytes = two 128-byte HBM lines per draw; 1 = iid)")
    ap.add_argument("--table-order", default="interleaved", choices=list(TABLE_ORDERS),
                    help="interleaved = ABNN_TABLE_DST_INTERLEAVED (8 adjacent destinations per 128-byte line), dst = ABNN_TABLE_DST_SORTED "
                         "(stable sort by destination neuron), given = generation order")
    ap.add_argument("--warm-frac", type=float, default=0.25,
                    help="fraction of neurons whose lastFired is pre-seeded inside the pre-spike window (SURVEY §8d 'warm' variant)")
    ap.add_argument("--src-view", default="snapshot", choices=["snapshot", "live"])
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer"], help="N>1: per-pass exchange (abnn_params.exchange)")
    ap.add_argument("--no-visits", action="store_true")
    ap.add_argument("--no-l2-persist", action="store_true")
    ap.add_argument("--structural", action="store_true", help="BASELINE configs[4]: prune + grow after every pass")
    ap.add_argument("--compact-every", type=int, default=64,
                    help="--structural: abnn_params.compact_every (1 = rebuild the table at every structural step; K > 1 = mark dead in "
                         "place + append behind the table, rebuild every K-th step: README.md:122-124 'compact periodically')")
    ap.add_argument("--cpu-syn", type=int, default=0, help="table of the CPU arm (default: the GPU arm's when the host has the memory, else 1e8)")
    ap.add_argument("--cpu-events", type=int, default=0, help="events per CPU step (default: sized for about 20-90 s in total)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-variants", action="store_true", help="no sampler / exact-mode sub-records, no parity record")
    a = ap.parse_args()
    if not a.syn:
        a.syn = 500_000_000 * max(1, a.gpus) if a.structural else 1_000_000_000
    if not a.events:
        a.events = 150_000_000 * max(1, a.gpus) // 8 if a.structural else 150_000_000
    if not a.hidden:
        a.hidden = 5_000_000 * max(1, a.gpus) // 8 if a.structural else 5_000_000
    return a


def kernel_source_hash():
    """Identity of the traversal kernel's source: an ncu capture is only quoted for the code it was taken from."""
    h = hashlib.sha256()
    for f in ("traversal.cu", "common.cuh"):
        h.update(open(os.path.join(ROOT, "abnn_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(workload_key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this workload and THIS kernel source
    (profiles/r2_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum), else None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        for e in t["captures"]:
            if e.get("workload_key") == workload_key and e.get("kernel_source_sha16") == kernel_source_hash():
                return e["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def peaks():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (10 ms period) that is
    started before the warm-up, so that even a 60 ms timed region (8 GPUs, 200 steps) holds samples; only the samples
    between mark_start() and mark_stop() are reported. Falls back to `nvidia-smi -lms` when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, period_s=0.01):
        self.index, self.period = index, period_s
        self.rows = []                      # (perf_counter time, sm MHz, max sm MHz, set of reasons)
        self.p = self.t = self.nv = None
        self.t0 = self.t1 = None
        self.halt = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = None
            try:                            # CUDA_VISIBLE_DEVICES may renumber the devices: find the NVML device by UUID
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
                try:
                    self.dev = pynvml.nvmlDeviceGetHandleByUUID(uuid)
                except TypeError:
                    self.dev = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.dev = None
            if self.dev is None:
                self.dev = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            # nvmlClocksEventReason* bit values (nvml.h)
            self.bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read_smi, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _poll_nvml(self):
        nv = self.nv
        while not self.halt.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev))
                self.rows.append((time.perf_counter(), sm, self.mx, {n for n, b in self.bits.items() if mask & b}))
            except Exception:
                pass
            self.halt.wait(self.period)

    def _read_smi(self):
        for line in self.p.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) < 9:
                continue
            try:
                self.rows.append((time.perf_counter(), float(r[1]), float(r[2]),
                                  {n for n, v in zip(self.NAMES, r[5:9]) if v.lower().startswith("active")}))
            except ValueError:
                continue

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_stop(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.t1 is None:
            self.mark_stop()
        self.halt.set()
        if self.p:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
        if self.t:
            self.t.join(timeout=2)
        if not self.p and not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"]}
        return self.summarise(self.rows, self.t0, self.t1, "nvml" if self.nv else "nvidia-smi")

    @staticmethod
    def summarise(rows, t0, t1, source):
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "source": source}
        inside = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        note = None
        if not inside:                      # timed region shorter than the sampling period: the sample closest to it
            mid = 0.5 * ((t0 if t0 is not None else rows[0][0]) + t1)
            inside = [min(rows, key=lambda r: abs(r[0] - mid))]
            note = "no sample fell inside the timed region; nearest sample, %.0f ms away" % (1e3 * abs(inside[0][0] - mid))
        reasons = set()
        for r in inside:
            reasons |= r[3]
        out = {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": float(max(r[2] for r in inside)),
               "reasons": sorted(reasons), "samples": len(inside), "source": source}
        if note:
            out["note"] = note
        return out


def stimulus_frames(n):
    from abnn_b200 import FunctionalDataset
    d = FunctionalDataset(N_IN, N_OUT)
    fin, fex = [], []
    for _ in range(n):
        fin.append(d.nextInput()); fex.append(d.nextExpected())
    return np.stack(fin), np.stack(fex)


def effective_order(args):
    """configs[4] runs on the dst-sorted order: the per-pass sorted insertion of grown synapses exists for that order."""
    return "dst" if args.structural and args.table_order == "interleaved" else args.table_order


def workload_params(args, p, rank, world, events, **over):
    """The benchmark's parameters on top of a defaults struct `p` (the library's abnn_default_params in the GPU arm, the
    oracle's independent copy in the CPU arm): north-star profile; the pre-spike window / refractory period are the
    reference's 5 / 2 PASSES (brain.metal:23-24) expressed in per-event ticks (one pass = `events` ticks)."""
    from abnn_b200 import capi                       # constants and struct layouts only; does not load the CUDA library
    orders = {"interleaved": capi.TABLE_DST_INTERLEAVED, "dst": capi.TABLE_DST_SORTED, "given": capi.TABLE_AS_GIVEN}
    vals = dict(n_input=N_IN, n_output=N_OUT, n_hidden=args.hidden, n_syn=args.syn, seed=42,
                sampler=capi.SAMPLER_PHILOX if args.sampler == "philox" else capi.SAMPLER_SWEEP,
                exec_mode=capi.EXEC_PARALLEL, window_pre=5 * events, refractory=2 * events,
                track_visits=0 if args.no_visits else 1, l2_persist=0 if args.no_l2_persist else 1,
                rank=rank, world_size=world, device=-1, sample_block=args.block, table_order=orders[effective_order(args)],
                src_view=capi.SRC_SNAPSHOT if args.src_view == "snapshot" else capi.SRC_LIVE,
                exchange=capi.EXCHANGE_PEER if args.exchange == "peer" else capi.EXCHANGE_NCCL)
    if args.structural:                              # configs[4]: Beta(2,8) weights, ~7 % below 0.05 at the start
        share = args.syn // world
        vals.update(w_prune=0.05, p_new=0.25, w_init=0.1, syn_capacity=share + (share >> 5) + (1 << 20), compact_every=args.compact_every)
    vals.update(over)
    for k, v in vals.items():
        setattr(p, k, v)
    return p


def warm_timestamps(n_neuron, frac, events, seed=7):
    """Pre-seed lastFired of `frac` of the neurons uniformly inside the pre-spike window so that the
    gated fraction is non-trivial from the first timed pass; clock starts after the window."""
    rng = np.random.default_rng(seed)
    lf = np.zeros(n_neuron, np.uint64)
    k = int(n_neuron * frac)
    idx = rng.choice(n_neuron, size=k, replace=False)
    start = 6 * events
    lf[idx] = rng.integers(start - 5 * events, start, size=k).astype(np.uint64)
    return lf, start


def host_memory_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2**30
    except Exception:
        return 0.0


def run_cpu(args, steps, warmup, budget_s):
    """The oracle port on all host threads: T dst-shards, one thread each (the multi-GPU partition). Every step is a pass
    of `ev` events (a bounded sample of the EVENTS_PER_PASS pass, same per-event work) over the CPU arm's table; `ev` is
    sized from a calibration pass so that steps + warmup passes take about budget_s. Returns a dict."""
    from abnn_b200 import capi
    from oracle import pyoracle as O
    T = os.cpu_count() or 1
    need_gb = 3.2 * 16e-9 * args.syn + 16            # table + its sort copy + slack
    syn = args.cpu_syn or (args.syn if host_memory_gb() > need_gb else 100_000_000)
    p = workload_params(args, O.default_params(capi.PROFILE_B200), 0, 1, args.events, n_syn=syn, exec_mode=capi.EXEC_SERIAL, syn_capacity=0)
    world = O.OracleWorld(p, T)
    t0 = time.perf_counter()
    th = [threading.Thread(target=s.init_graph, args=(capi.GRAPH_ER_BETA, 1)) for s in world.shards]
    [t.start() for t in th]; [t.join() for t in th]
    world._sync_counts()
    t_init = time.perf_counter() - t0
    n_neuron = N_IN + N_OUT + args.hidden
    lf, start = warm_timestamps(n_neuron, args.warm_frac, args.events)
    for s in world.shards:
        s.upload_timestamps(lf, None); s.clock = start; s.set_reward(0.01)
    fin, fex = stimulus_frames(steps + warmup + 1)
    ev = args.cpu_events
    if not ev:                                       # calibration: 2M events, then size the step
        t0 = time.perf_counter(); world.run_pass(2_000_000); rate = 2_000_000 / (time.perf_counter() - t0)
        ev = int(min(args.events, max(1_000_000, rate * budget_s / max(1, steps + warmup))))
    times, gated, evs = [], 0, 0
    for it in range(steps + warmup):
        t0 = time.perf_counter()
        for s in world.shards:
            s.inject_inputs(fin[it], 1000.0); s.teacher_force(fex[it], float(it & 1))
        st = world.run_pass(ev)
        world.shards[0].readout_filtered(fex[it])
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt); gated += st.gated; evs += st.events
    ms = 1e3 * float(np.mean(times))
    # the 1-thread figure SURVEY.md §8d asks for: shard 0 alone executes its share of one more pass
    t0 = time.perf_counter()
    st1 = world.shards[0].run_pass(ev)
    one = {"value": st1.events / (time.perf_counter() - t0), "sample": f"{st1.events:,} events of shard 0 ({syn // T:,} synapses), one thread"}
    same = syn == args.syn
    sample = (f"{steps} steps x {ev:,} events (a bounded sample of the {args.events:,}-event pass, same per-event work) on a {syn:,}-synapse / "
              f"{n_neuron:,}-neuron ER-Beta graph ({syn * 16 / 1e9:.1f} GB table" + ("" if same else f" instead of {args.syn * 16 / 1e9:.1f} GB") +
              f"), {T} dst-shards on {T} threads; table built in {t_init:.1f} s")
    return {"value": ev / (ms * 1e-3), "ms": ms, "cores": T, "sample": sample, "gated_fraction": gated / max(1, evs), "one_thread": one,
            "events_per_step": ev, "syn": syn, "steps": steps, "warmup": warmup, "same_table": same}


def config_dict(args, world, exec_mode="parallel", **over):
    c = {"sampler": args.sampler, "sample_block": args.block, "table_order": effective_order(args),
         "exec_mode": exec_mode, "clock": "per_event", "graph": "ER endpoints, Beta(2,8) weights (Philox)", "window_pre_passes": 5,
         "refractory_passes": 2, "warm_fraction": args.warm_frac, "track_visits": not args.no_visits,
         "parallelism": f"dst-shard x{world}" if world > 1 else "single GPU", "l2": "inputs larger than L2 (16 GB table, random gathers)"}
    c.update(over)
    return c


def make_brain(args, rank, world, local_rank, dist, **over):
    """A handle on this rank's GPU with the workload's graph, warm timestamps and reward."""
    import torch
    from abnn_b200 import Brain, capi
    p = workload_params(args, capi.default_params(capi.PROFILE_B200), rank, world, args.events, **over)
    p.device = local_rank
    b = Brain(p)
    if world > 1:
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            import ctypes as C
            raw = C.create_string_buffer(128)
            capi.check(b.lib.abnn_comm_unique_id(raw), "abnn_comm_unique_id")
            idbuf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
        idbuf = idbuf.cuda()
        dist.broadcast(idbuf, 0)
        b.comm_init(bytes(idbuf.cpu().numpy().tobytes()))
    b.init_graph(capi.GRAPH_ER_BETA, 1)
    n_neuron = N_IN + N_OUT + args.hidden
    lf, start = warm_timestamps(n_neuron, args.warm_frac, args.events)
    b.upload_timestamps(lf, None)
    b.clock = start
    b.set_reward(0.01)
    return b


def sub_record(args, name, peak, passes=5, **over):
    """One sub-record beside the headline (single GPU): the same workload with `over` changed, a few passes through
    abnn_run_pass, traversal time from the handle's CUDA events."""
    b = make_brain(args, 0, 1, 0, None, **over)
    try:
        tm, pm, gated, fired, evs = [], [], 0, 0, 0
        for it in range(passes + 1):
            st = b.run_pass(args.events)
            if it:                                    # the first pass warms lazily configured kernels / allocations
                tm.append(st.traverse_ms); pm.append(st.device_ms); gated += st.gated; fired += st.fired; evs += st.events
        g = gated / max(1, evs)
        k_ms, p_ms = float(np.mean(tm)), float(np.mean(pm))
        return {"value": args.events / (p_ms * 1e-3), "unit": "events/s", "pass_ms": p_ms, "kernel_ms": k_ms, "gated_fraction": g,
                "fire_fraction": fired / max(1, evs), "roofline_frac": args.events * (16.0 + 16.0 * g) / (k_ms * 1e-3) / 1e9 / peak,
                "passes": passes, "what": name}
    finally:
        b.close()


def parity_record(args, rank, world, local_rank, dist):
    """N>1, after the timed region, on fresh handles of the same workload: (1) after 3 PARALLEL passes every rank holds
    the SAME gate words and they equal slack_word(clock, lastFired, window_pre) of the downloaded timestamps; (2) an
    EXACT-mode run of the same passes (bit-identical to the serial oracle, tests/) gives the same lastVisited (an
    order-free max over the sampled events) and the same first-pass candidate count, gated / fired within 2 %."""
    from abnn_b200 import capi
    passes, out, res = 3, {}, {}
    for mode, name in ((capi.EXEC_PARALLEL, "parallel"), (capi.EXEC_EXACT, "exact")):
        b = make_brain(args, rank, world, local_rank, dist, exec_mode=mode)
        stats = [b.run_pass(args.events) for _ in range(passes)]
        lf, lv = b.timestamps()
        info = b.info()
        words = valid = None
        if mode == capi.EXEC_PARALLEL:
            words, valid = b.gate_words()
        res[name] = (stats, lf, lv, info, words, valid, b.clock)
        b.close()
    (sp, lfp, lvp, info, words, valid, clock), (se, lfe, lve, _, _, _, _) = res["parallel"], res["exact"]
    lo, hi = int(info.neuron_lo), int(info.neuron_hi)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
    # (1) gate words: identical on every rank, equal to the closed form of the exchanged lastFired (common.cuh:slack_word)
    clk, win = np.uint64(clock), np.uint64(5 * args.events)
    age = clk - np.minimum(lfp, clk)
    room = win - np.minimum(age, win)
    want = np.where(lfp > clk, np.uint64(0xFFFFFFFF), np.where(age > win, np.uint64(0), np.minimum(room, np.uint64(0xFFFFFFFD)) + np.uint64(1))).astype(np.uint32)
    mine = {"gate_sha": sha(words), "gate_valid": bool(valid), "gate_equals_slack_word": bool(np.array_equal(words, want)),
            "lastFired_sha": sha(lfp),
            "lastVisited_owned_equal_exact": bool(np.array_equal(lvp[lo:hi], lve[lo:hi])),
            "candidates_pass0": [int(sp[0].candidates), int(se[0].candidates)],
            "gated": [int(sum(s.gated for s in sp)), int(sum(s.gated for s in se))],
            "fired": [int(sum(s.fired for s in sp)), int(sum(s.fired for s in se))]}
    allr = [None] * world
    dist.all_gather_object(allr, mine)
    if rank == 0:
        gp, ge = sum(r["gated"][0] for r in allr), sum(r["gated"][1] for r in allr)
        fp, fe = sum(r["fired"][0] for r in allr), sum(r["fired"][1] for r in allr)
        out = {"passes": passes, "ranks": world,
               "gate_words_identical_on_all_ranks": len({r["gate_sha"] for r in allr}) == 1 and all(r["gate_valid"] for r in allr),
               "gate_words_equal_slack_word_of_lastFired": all(r["gate_equals_slack_word"] for r in allr),
               "lastFired_identical_on_all_ranks": len({r["lastFired_sha"] for r in allr}) == 1,
               "lastVisited_equals_exact_mode": all(r["lastVisited_owned_equal_exact"] for r in allr),
               "candidates_pass0_equal_exact_mode": all(r["candidates_pass0"][0] == r["candidates_pass0"][1] for r in allr),
               "gated_parallel_vs_exact": [gp, ge], "fired_parallel_vs_exact": [fp, fe],
               "gated_within_2pct": abs(gp - ge) <= 0.02 * ge, "fired_within_2pct": abs(fp - fe) <= 0.02 * fe + 100}
        out["ok"] = all(v for v in out.values() if isinstance(v, bool))
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = args.steps, max(args.warmup, 0)
    shape = "BASELINE configs[4] (structural plasticity every pass)" if args.structural else "constants.h shape"
    workload = (f"{shape}: {args.hidden:,} hidden + {N_IN} in + {N_OUT} out, {args.syn:,} synapses "
                f"({args.syn * 16 / 1e9:.1f} GB SynapsePacked), {args.events:,} events/pass")

    if args.impl == "reference":
        if rank != 0:
            return
        r = run_cpu(args, max(1, K), W, budget_s=90.0)
        wl = (f"{shape}: {args.hidden:,} hidden + {N_IN} in + {N_OUT} out, {r['syn']:,} synapses ({r['syn'] * 16 / 1e9:.1f} GB SynapsePacked), "
              f"steps of {r['events_per_step']:,} events sampled from the {args.events:,}-event pass")
        print(json.dumps({
            "impl": "reference", "metric": "synaptic events/sec", "value": r["value"], "unit": "events/s", "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms"], "higher_is_better": True,
            "scaling": "weak" if args.structural else "strong", "vs_baseline": None, "dtype": "u64 timestamps, f32 weights", "data": "synthetic",
            "config": dict(config_dict(args, 1, "serial per dst-shard (oracle), one shard per host thread",
                                       parallelism=f"host threads x{r['cores']} (one dst-shard each), no GPU",
                                       l2=f"{r['syn'] * 16 / 1e9:.1f} GB table in host memory, random gathers"), workload=wl,
                           same_table_as_gpu_arm=r["same_table"], events_per_step=r["events_per_step"]),
            "cpu_baseline": {"value": r["value"], "unit": "events/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
                             "one_thread": r["one_thread"], "gated_fraction": r["gated_fraction"]},
            "e2e": {"value": r["value"], "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference has no CPU traversal and its Metal app cannot be built on Linux; this is the oracle port of its algorithm "
                    "(oracle/oracle_b.cpp); events/s is per-event work, so a sampled step measures the same rate as a full pass"}))
        return

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    b = make_brain(args, rank, world, local_rank, dist)
    n_neuron = N_IN + N_OUT + args.hidden
    info = b.info()
    fin, fex = stimulus_frames(2 * (K + W) + 16)
    pin_in = torch.from_numpy(fin).pin_memory().numpy()
    pin_ex = torch.from_numpy(fex).pin_memory().numpy()

    def barrier():
        b.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    struct_ms, struct_stats = [], []

    def step_device(it, rates=False):
        # BrainEngine::run_one_pass as one C-ABI call: stage frame, inject, teacher forcing, traversal, exchange,
        # read-out step (replayed from a CUDA graph after the second call); configs[4]: + the structural step
        r = b.engine_step(pin_in[it], pin_ex[it], 1000.0, float(it & 1), args.events, want_rates=rates)
        if args.structural:
            t0 = time.perf_counter()
            s = b.prune_and_grow()                   # synchronises
            struct_ms.append(1e3 * (time.perf_counter() - t0)); struct_stats.append(s)
        return r

    # ---- value: device-timed, K steps ------------------------------------------------------------
    sampler = ClockSampler(local_rank)       # polls from before the warm-up; reports the samples of the timed region
    if rank == 0:
        sampler.start()
    for it in range(W):
        step_device(it)
    barrier()
    del struct_ms[:], struct_stats[:]
    if rank == 0:
        sampler.mark_start()
    b.timer_mark(0)
    for it in range(W, W + K):
        step_device(it)
    b.timer_mark(1)
    barrier()
    if rank == 0:
        sampler.mark_stop()
    ms_total = b.timer_elapsed(0, 1)
    clocks = sampler.stop() if rank == 0 else None
    n_after = int(b.info().n_syn_local)
    struct_mean = float(np.mean(struct_ms)) if struct_ms else 0.0
    struct_last = struct_stats[-1] if struct_stats else None

    # ---- roofline: traversal kernel alone, per launch (CUDA events around the kernel) ---------------
    trav_ms, pass_ms, gated, fired, evs = [], [], 0, 0, 0
    for it in range(min(K, 10)):
        b.inject_inputs(pin_in[W + K + it], 1000.0)
        b.teacher_force(pin_ex[W + K + it], float(it & 1))
        st = b.run_pass(args.events)
        trav_ms.append(st.traverse_ms); pass_ms.append(st.device_ms); gated += st.gated; fired += st.fired; evs += st.events
    barrier()

    # ---- e2e: host buffers in, filtered read-out back to the host, every step -----------------------
    t0 = time.perf_counter()
    for it in range(K):
        j = (W + K + 10 + it) % len(pin_in)
        step_device(j, rates=True)                   # H2D frame, D2H rates, sync
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- read-dominated regime (the reference's steady state, g -> 0): no neuron has fired inside
    # the window, no input spikes; every event still samples, gathers its record, reads the gate word and
    # marks lastVisited. Reported beside the headline, never instead of it.
    quiet_ms, quiet_g = [], 0
    b.upload_timestamps(np.zeros(n_neuron, np.uint64), None)
    b.clock = 1000 * args.events
    for it in range(6):
        st = b.run_pass(args.events)
        if it >= 2:
            quiet_ms.append(st.traverse_ms); quiet_g += st.gated
    barrier()
    quiet_mean = float(np.mean(quiet_ms))
    b.close()

    if world > 1:
        q = torch.tensor([quiet_mean], dtype=torch.float64, device="cuda")
        dist.all_reduce(q, op=dist.ReduceOp.MAX)
        quiet_mean = float(q.item())
        t = torch.tensor([ms_total, e2e_s, float(np.mean(trav_ms)), float(np.mean(pass_ms)), struct_mean], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {"traverse_ms": float(np.mean(trav_ms)), "gated_fraction": gated / max(1, evs), "events_per_pass": evs / max(1, min(K, 10))})
        cnt = torch.tensor([gated, fired, evs, n_after], dtype=torch.float64, device="cuda")
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        ms_total, e2e_s, trav_mean, pass_mean, struct_mean = (float(x) for x in t.tolist())
        gated, fired, evs, n_after = (float(x) for x in cnt.tolist())
    else:
        trav_mean, pass_mean = float(np.mean(trav_ms)), float(np.mean(pass_ms))
        per_rank = None

    parity = None
    if world > 1 and not args.skip_variants and not args.structural:
        parity = parity_record(args, rank, world, local_rank, dist)

    if rank == 0:
        ms_step = ms_total / K
        value = args.events / (ms_step * 1e-3)
        g = gated / max(1.0, evs)
        b_alg = 16.0 + 16.0 * g
        peak, peak_kind = peaks()
        # per-GPU roofline of the traversal kernel: this rank's events per launch
        ev_per_launch = evs / (min(K, 10) * world)
        wkey = f"{args.hidden}/{args.syn}/{args.events}/{args.sampler}/b{args.block}/{effective_order(args)}/{args.src_view}/w{world}"
        achieved = ev_per_launch * b_alg / (trav_mean * 1e-3) / 1e9
        traffic = ncu_traffic(wkey)
        line32 = args.sampler == "philox" and args.block in (1, 8, 16) and args.src_view == "snapshot"
        line = {
            "metric": "synaptic events/sec", "value": value, "unit": "events/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak" if args.structural else "strong", "vs_baseline": None,
            "dtype": "u64 timestamps, f32 weights", "data": "synthetic",
            "config": dict(config_dict(args, world), workload=workload, l2_persist_bytes=int(info.l2_persist_bytes),
                           exchange=args.exchange if world > 1 else None),
            "gated_fraction": g, "fire_fraction": fired / max(1.0, evs),
            "read_dominated_regime": {"value": args.events / (quiet_mean * 1e-3), "unit": "events/s", "kernel_ms": quiet_mean,
                                      "gated_fraction": quiet_g / (4.0 * args.events),
                                      "roofline_frac": args.events / world * 16.0 / (quiet_mean * 1e-3) / 1e9 / peak,
                                      "note": "same table, no neuron inside the pre-spike window (B_alg = 16 B/event)"},
            # where a step goes (device time, max over ranks): traversal kernel | + word build, end-of-pass, fold, exchange
            # | + inject, teacher, read-out and launch gaps = ms_per_step
            "step_breakdown_ms": {"traverse": trav_mean, "pass_with_exchange": pass_mean, "step": ms_step, "per_rank": per_rank},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_kind": peak_kind,
                         "kernel": "k_traverse_line32" if line32 else ("k_traverse_line" if args.block == 8 else "k_traverse_parallel"),
                         "kernel_ms": trav_mean, "alg_bytes_per_event": b_alg, "alg_bytes_per_launch": ev_per_launch * b_alg,
                         "dram_frac": (traffic / (trav_mean * 1e-3) / 1e9 / peak) if traffic else None,
                         "kernel_source_sha16": kernel_source_hash()},
            "e2e": {"value": args.events * K / e2e_s, "unit": "events/s",
                    "h2d_bytes_per_step": int((N_IN + N_OUT + 2) * 4), "d2h_bytes_per_step": int(N_OUT * 4)},
            # per step (abnn_engine_step; one CUDA-graph launch when captured): k_step_prologue (inject + teacher forcing + head
            # words), traversal kernel, k_end_pass, k_fold_prepare32 (fold + next pass's words), k_readout; the NCCL kernels of the
            # exchange and (--structural) the kernels of the structural step are not counted
            "gpu_launches": 5 * K,
            "clocks": clocks,
        }
        if args.structural:
            line["structural"] = {"ms_per_structural_step": struct_mean, "fraction_of_step": struct_mean / ms_step if ms_step else None,
                                  "compact_every": args.compact_every, "ms_per_structural_step_max": float(np.max(struct_ms)) if struct_ms else None,
                                  "n_syn_after": n_after, "pruned_last": int(struct_last.pruned) if struct_last else None,
                                  "appended_last": int(struct_last.appended) if struct_last else None,
                                  # stable compaction + sorted insertion out of place: 16 B read (count) + 16 B read + 16 B written per record
                                  "sweep_roofline_frac": ((48.0 * n_after / world) / (struct_mean * 1e-3) / 1e9 / peak
                                                          if struct_mean and args.compact_every <= 1 else None),
                                  # the host-timed call starts while the pass is still running on the device: without the traversal
                                  "ms_per_structural_step_excl_traversal": struct_mean - trav_mean,
                                  "note": "ms_per_step includes the structural step (the device timer spans the synchronising abnn_prune_and_grow); "
                                          "use --steps >= 2 * compact_every so that the timed region holds its share of rebuild steps"}
        if parity is not None:
            line["parity"] = parity
        if world == 1 and not args.skip_variants and not args.structural:
            from abnn_b200 import capi
            subs = {}
            for name, over in (("iid_sampler", dict(sample_block=1, table_order=capi.TABLE_AS_GIVEN)),
                               ("line8_dst_sorted", dict(sample_block=8, table_order=capi.TABLE_DST_SORTED)),
                               ("line8_interleaved", dict(sample_block=8, table_order=capi.TABLE_DST_INTERLEAVED))):
                try:
                    subs[name] = sub_record(args, name, peak, **over)
                except Exception as e:                # a sub-record must never take the headline down
                    subs[name] = {"error": str(e)[:200]}
            subs["iid_sampler"]["note"] = ("SURVEY §8.0 sampler edge(e) = mulhi(philox(seed,e), N_SYN), table in generation order, on "
                                           "k_traverse_line32<.,.,1> (one 16-byte cp.async per event); bounded by B200's 128-byte DRAM fetch per "
                                           "random 16-byte gather (36.7e9 bare gathers/s measured, profiles/r1_notes.md §1)")
            subs["line8_dst_sorted"]["note"] = ("round-1 headline layout: bursts of 8 events per neuron lower the fire rate by a few per cent "
                                                "(tests/test_gpu_equivalence.py)")
            line["samplers"] = subs
            try:
                line["exact_mode"] = sub_record(args, "EXACT execution (conflict-free, bit-identical to the serial order)", peak, passes=3,
                                                exec_mode=capi.EXEC_EXACT)
            except Exception as e:
                line["exact_mode"] = {"error": str(e)[:200]}
        if world == 1 and not args.skip_cpu:
            r = run_cpu(args, 3, 1, budget_s=20.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": "events/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
                                    "one_thread": r["one_thread"], "gated_fraction": r["gated_fraction"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
