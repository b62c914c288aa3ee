/* include/abnn.h — C-ABI of the B200-native ABNN traversal engine (libabnn_b200.so).
 *
 * This is the drop-in boundary for the reference's `Brain` class (the only place the reference
 * touches its GPU backend): every entry point below names the reference interface it replaces
 * (paths relative to the reference repo root). Plain C: pointers and sizes only, no C++/torch
 * types, never throws. All functions return 0 on success or a negative abnn_status; the message
 * for the last failure on the calling thread is abnn_last_error().
 *
 * Ownership: the handle owns all device memory, streams, events and communicators. The caller
 * owns every host pointer it passes; no pointer handed out by the library outlives the handle.
 * Threading: a handle is single-caller (externally synchronised), like the reference's one worker
 * thread (abnn/src/core/brain-engine.cpp:196-200). Work is enqueued on the handle's stream and is
 * asynchronous unless the entry point says it synchronises.
 *
 * There is no CPU fallback: abnn_create fails with ABNN_ERR_NO_DEVICE when no CUDA device exists.
 */
#ifndef ABNN_H_
#define ABNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABNN_ABI_VERSION 1u

/* ---- status codes ---------------------------------------------------------------------- */
enum abnn_status {
    ABNN_OK              =  0,
    ABNN_ERR_INVALID     = -1,  /* bad argument / params                                     */
    ABNN_ERR_NO_DEVICE   = -2,  /* no CUDA device (no CPU fallback exists)                   */
    ABNN_ERR_CUDA        = -3,  /* CUDA runtime failure, see abnn_last_error()               */
    ABNN_ERR_IO          = -4,  /* file could not be opened / short read / short write       */
    ABNN_ERR_SHAPE       = -5,  /* .bnn header does not match the handle (brain.cpp:174)     */
    ABNN_ERR_CAPACITY    = -6,  /* synapse table / staging capacity exceeded                 */
    ABNN_ERR_COMM        = -7,  /* NCCL failure or communicator missing                      */
    ABNN_ERR_UNSUPPORTED = -8   /* mode combination not available                            */
};

/* ---- T1: the synapse record (abnn/src/core/brain/brain.h:21, kernels/brain.metal:11) ----- */
typedef struct abnn_synapse {
    uint32_t src, dst;
    float    w, pad;
} abnn_synapse;                     /* 16 bytes, 16-byte aligned in device memory */

/* ---- modes ------------------------------------------------------------------------------ */
enum abnn_sampler {          /* which synapse event e touches                                 */
    ABNN_SAMPLER_SWEEP  = 0, /*  edge = e (thread t processes synapse t: brain.metal:60-61,70) */
    ABNN_SAMPLER_PHILOX = 1  /*  edge = mulhi64(philox(seed,e).xy, n_syn)  (README.md:77)     */
};
enum abnn_release_rng {      /* the uniform draw of the release test                          */
    ABNN_RNG_XORSHIFT = 0,   /*  rand01(tid ^ now)             (brain.metal:15-19,92)         */
    ABNN_RNG_PHILOX   = 1    /*  (philox(seed,e).z >> 8) * 2^-24                              */
};
enum abnn_clock_mode {
    ABNN_CLOCK_PER_PASS  = 0,/*  every event of pass c sees now = c; clock = c+1 after the pass
                                 ("hold-clock" reading of brain.metal:64-68,129)              */
    ABNN_CLOCK_PER_EVENT = 1 /*  now = clock + e; clock += events after the pass (README.md:62-63,85) */
};
enum abnn_exec_mode {
    ABNN_EXEC_SERIAL   = 0,  /*  one device thread walks the events in order (bit-exact oracle order) */
    ABNN_EXEC_EXACT    = 1,  /*  conflict-free parallel: bit-identical to SERIAL (needs SNAPSHOT src view,
                                 PASS_STEP r-bar)                                             */
    ABNN_EXEC_PARALLEL = 2   /*  fully parallel batches; same-dst conflicts resolved with warp
                                 match + 64-bit atomicMax; statistical parity                 */
};
enum abnn_src_view {
    ABNN_SRC_LIVE     = 0,   /*  lastFired[src] read live (brain.metal:73) — single GPU only  */
    ABNN_SRC_SNAPSHOT = 1    /*  lastFired[src] read from the pass-start snapshot (GPU-count invariant) */
};
enum abnn_rbar_mode {
    ABNN_RBAR_PASS_STEP  = 0,/*  r-bar += alpha*(R - r-bar) once per pass, after the events    */
    ABNN_RBAR_METAL_TID0 = 1 /*  event 0 updates r-bar iff it passes every gate (brain.metal:110-113); SERIAL only */
};
enum abnn_graph_kind {
    ABNN_GRAPH_REFERENCE = 0,/*  build_random_graph (brain-engine.cpp:31-53): mt19937(seed), dense in->out
                                 U[.4,.8) then hid->hid U[.1,.2)                               */
    ABNN_GRAPH_ER_BETA   = 1 /*  Erdos-Renyi endpoints, w ~ Beta(2,8) (README.md:134-135), Philox-keyed by edge */
};
enum abnn_table_order {     /* HBM layout of this rank's synapse table                       */
    ABNN_TABLE_AS_GIVEN   = 0,/*  records stay in upload / generation order (brain-engine.cpp:37-52)  */
    ABNN_TABLE_DST_SORTED = 1,/*  stable sort by dst after every upload / init / load / growth step: the
                                 records that target one neuron are contiguous, so a 128-byte line of
                                 the table (sample_block = 8) touches one lastFired/lastVisited sector
                                 instead of eight. Table order is part of the semantics (edge(e) indexes
                                 it); the oracle applies the same stable sort.                  */
    ABNN_TABLE_DST_INTERLEAVED = 2 /* DST_SORTED with the records of every group of 8 consecutive neurons (id >> 3; 16
                                 neurons when sample_block >= 16, so that a 256-byte sample group spans 16 neurons)
                                 interleaved: inside a group the order is (rank of the record among its
                                 destination's records, destination), i.e. row r of a group holds the r-th record
                                 of each of its neurons. A 128-byte line still touches ONE sector of the per-neuron
                                 arrays (8 adjacent neurons), but no two events of a sample group hit the same
                                 neuron: every neuron sees the event arrival statistics of the iid sampler (with
                                 DST_SORTED its events arrive in bursts of sample_block, which lowers the fire rate
                                 by a few per cent — DESIGN.md §2). Re-derived after every upload / init / load
                                 and after every structural step that changed the table.        */
};
enum abnn_exchange {        /* per-pass timestamp exchange of a sharded handle (world_size > 1)          */
    ABNN_EXCHANGE_NCCL = 0,   /*  ncclAllGather of the owned slices                       */
    ABNN_EXCHANGE_PEER = 1    /*  each rank stores its slice of gate words straight into every peer's array over NVLink
                                 (CUDA-IPC mappings, flag rounds instead of a collective; falls back to NCCL on every
                                 rank if any rank cannot map a peer). Needs all ranks on one NVLink/NVSwitch node. */
};
enum abnn_profile {
    ABNN_PROFILE_METAL_PARITY = 0, /* SWEEP, XORSHIFT, PER_PASS, SERIAL, LIVE, METAL_TID0, budget 2560 */
    ABNN_PROFILE_NORTH_STAR   = 1, /* PHILOX, PHILOX, PER_EVENT, PARALLEL, SNAPSHOT, PASS_STEP          */
    ABNN_PROFILE_B200         = 2  /* NORTH_STAR with the layout the B200 kernel is built for: sample_block 16 (256 bytes
                                      = two 128-byte lines per draw, the size at which random HBM3e reads reach the copy
                                      bandwidth) over a DST_INTERLEAVED table — what bench.py times                  */
};

/* ---- L3: every compile-time knob of the reference as a runtime parameter ------------------
 * (abnn/src/core/constants.h:2-19, kernels/brain.metal:22-31, brain/brain.h:17-19)            */
typedef struct abnn_params {
    uint32_t struct_size;          /* = sizeof(abnn_params); ABI check                          */
    uint32_t abi_version;          /* = ABNN_ABI_VERSION                                        */

    /* shape (Brain ctor, brain.cpp:21-27); neuron ids: [0,n_input) inputs, then outputs, then hidden */
    uint32_t n_input;              /* NUM_INPUTS  256                                           */
    uint32_t n_output;             /* NUM_OUTPUTS 256                                           */
    uint64_t n_hidden;             /* NUM_HIDDEN  5'000'000                                     */
    uint64_t n_syn;                /* NUM_SYN     1'000'000'000 (global, all ranks)             */
    uint64_t syn_capacity;         /* records this rank can hold; 0 = its share of n_syn        */

    uint64_t seed;                 /* Philox key for events / inject / teacher / growth         */

    uint32_t sampler;              /* abnn_sampler                                              */
    uint32_t release_rng;          /* abnn_release_rng                                          */
    uint32_t clock_mode;           /* abnn_clock_mode                                           */
    uint32_t exec_mode;            /* abnn_exec_mode                                            */
    uint32_t src_view;             /* abnn_src_view                                             */
    uint32_t rbar_mode;            /* abnn_rbar_mode                                            */
    uint32_t max_spikes_per_pass;  /* kMaxSpikes 2560 (brain.h:18); 0 = unlimited. Saturating.   */
    uint32_t track_visits;         /* 1: lastVisited[dst] = max(.,now) on every event (README.md:84);
                                      0: never written (what brain.metal does, :44)             */

    uint64_t window_pre;           /* WINDOW_PRE 5  : gate  now - lastF[src] >  window_pre -> skip */
    uint64_t refractory;           /* REFRACTORY 2  : gate  now - lastF[dst] <= refractory -> skip */
    uint64_t teacher_gap;          /* teacher spike only if now - lastF[out] > teacher_gap (1; brain-engine.cpp:130) */

    float base_scale;              /* BASE_SCALE 0.8   p = clamp(w*w*base_scale,0,1)            */
    float a_ltp;                   /* _aLTP 0.04                                                */
    float a_ltd;                   /* _aLTD 0.02                                                */
    float w_min;                   /* _wMin 0.001                                               */
    float w_max;                   /* _wMax 1.0                                                 */
    float eta_home;                /* ETA_HOME 1e-6                                             */
    float target_rate_hz;          /* TARGET_RATE_HZ 1000                                       */
    float home_tick_hz;            /* 1e6: est rate = home_tick_hz / isi (brain.metal:117)      */
    float eta_reward;              /* ETA_REWARD 1e-3                                           */
    float alpha_rbar;              /* ALPHA_RBAR 1e-3                                           */

    /* structural plasticity (README.md:120-127; absent from the reference code)                 */
    float w_prune;                 /* prune iff w < w_prune; 0 = never                          */
    float p_new;                   /* on fire: grow iff philox(seed,e).w * 2^-32 < p_new; 0 = never */
    float w_init;                  /* weight of a grown synapse                                 */

    /* read-out (brain-engine.cpp:145-186, output-filter/rate-filter.h:22-59)                    */
    float    rate_alpha;           /* 0.5  spike-rate EMA                                       */
    float    peak_decay;           /* PEAK_DECAY 0.999                                          */
    float    peak_init;            /* maxObserved initial 0.5 (brain-engine.h:54)               */
    uint32_t use_fir;              /* USE_FIR true                                              */
    uint32_t fir_size;             /* 20 (rate-filter.h:14); <= ABNN_MAX_FIR                    */
    uint32_t reward_window;        /* WIN_SIZE_ 1000 passes (brain-engine.h:81)                 */
    double   filter_tau;           /* FILTER_TAU 0.02                                           */
    double   dt_sec;               /* dT_SEC 0.0009                                             */
    double   loss0;                /* lastLoss_ initial 0.25 (brain-engine.h:83)                */

    /* placement: one handle per GPU; neurons (and the synapses that target them) are sharded by
     * destination neuron over world_size ranks                                                  */
    int32_t  device;               /* CUDA device ordinal; -1 = current device                  */
    uint32_t rank;                 /* 0 .. world_size-1                                         */
    uint32_t world_size;           /* 1 = single GPU                                            */
    uint32_t l2_persist;           /* 1 (PARALLEL execution only): keep the per-neuron 32-bit arrays resident in L2: raises the
                                      DEVICE-wide cudaLimitPersistingL2CacheSize to what they need and attaches an
                                      access-policy window to the handle's stream; the limit the device had before is put
                                      back when the last such handle of the process is destroyed. 0, or SERIAL / EXACT
                                      execution: no device-wide state is touched */

    /* PHILOX sampler granularity: events are drawn in groups of sample_block consecutive events that
     * process sample_block consecutive table records starting at a Philox-chosen, block-aligned
     * position: edge(i) = B*mulhi64(q.xy, ceil(n/B)) + i%B  (skipped if >= n), q = philox(seed, i - i%B): ONE Philox
     * call per group; event k = i%B of the group takes its release draw from fmix32(q.z + k*0x9E3779B9) and its
     * synaptogenesis trial from fmix32(q.w + k*0x85EBCA6B) (fmix32 = MurmurHash3's 32-bit finaliser; B = 1: q.z, q.w).
     * 1 = every event draws its own edge (README.md:77). 8 = one 128-byte HBM line per draw (B200's DRAM fetch
     * granularity: a random 16-byte gather costs a whole line, profiles/r1_notes.md §1). 16 = 256 bytes per draw, the
     * size at which random HBM3e reads reach the copy bandwidth (profiles/r2_notes.md §1).
     * Power of two, <= 32. Every edge is still sampled with equal probability, but with sample_block > 1 the events of a
     * group arrive together: over a DST_SORTED table that is a burst on ONE neuron (fire rate 2-3 % below iid sampling),
     * over a DST_INTERLEAVED table the group's events hit different neurons and the statistics are those of the iid
     * sampler (tests/test_gpu_equivalence.py). */
    uint32_t sample_block;
    uint32_t table_order;          /* abnn_table_order                                          */
    uint32_t prune_in_place;       /* 0: pruning compacts into a second table when the device has memory for one (count +
                                      scatter passes, every tile independent), else in place; 1: always in place (single-
                                      pass chained scan: no second table, about 1.7x slower)                          */
    uint32_t exchange;             /* abnn_exchange: how sharded PARALLEL runs publish their gate words after a pass    */
    uint32_t compact_every;        /* structural steps between two rebuilds of the table. 0 / 1: every abnn_prune_and_grow is a
                                      stable prune-compaction + ordered (sorted) insertion. K > 1 — README.md:122-124 "remove,
                                      compact PERIODICALLY": steps 0, K, 2K, ... (counted since the table was uploaded /
                                      initialised / loaded) rebuild the table like that; the steps in between touch only what
                                      changed: a record whose weight fell below w_prune is marked dead IN PLACE (src =
                                      0xFFFFFFFF; its slot stays, an event that samples it does nothing) and the grown synapses
                                      are appended behind the table in tick order, to be merged into their destination's run by
                                      the next rebuild. Needs w_init >= w_prune.                                              */
    uint32_t reserved_;
} abnn_params;

#define ABNN_DEAD_SRC 0xFFFFFFFFu   /* abnn_synapse.src of a pruned record that waits for the next rebuild (compact_every > 1) */

#define ABNN_MAX_FIR 64u

typedef struct abnn_info {
    uint32_t n_input, n_output;
    uint64_t n_hidden, n_neuron;          /* Brain::n_input..n_neuron (brain.h:48-51)           */
    uint64_t n_syn_global;                /* Brain::n_syn (brain.h:52), summed over ranks at creation/last structural step */
    uint64_t n_syn_local, syn_capacity;   /* this rank's live records / capacity                */
    uint64_t neuron_lo, neuron_hi;        /* destination-neuron range this rank owns            */
    uint64_t neuron_slice;                /* ceil(n_neuron / world_size)                        */
    uint32_t rank, world_size;
    int32_t  device;
    uint32_t sm_count;
    uint64_t l2_bytes, l2_persist_bytes;  /* device L2 size / bytes actually set aside          */
    uint64_t pass_index, clock, event_base;
} abnn_info;

typedef struct abnn_pass_stats {
    uint64_t events;        /* events this rank executed in the pass                            */
    uint64_t gated;         /* events that passed window + refractory + budget gates (weight written) */
    uint64_t fired;         /* events that fired (timestamp written)                            */
    uint64_t candidates;    /* events that passed the src (pre-spike) window                      */
    uint64_t grown;         /* synaptogenesis candidates staged this pass                       */
    uint64_t clock;         /* clock after the pass                                             */
    double   device_ms;     /* device time of the pass (CUDA events), incl. timestamp exchange  */
    double   traverse_ms;   /* device time of the traversal kernel(s) alone                     */
} abnn_pass_stats;

typedef struct abnn_structural_stats {
    uint64_t n_before, pruned, appended, n_after;   /* local counts                             */
    uint64_t dropped;       /* growth candidates dropped because capacity was reached           */
} abnn_structural_stats;

typedef struct abnn_handle abnn_handle;

/* ---- library ---------------------------------------------------------------------------- */
const char* abnn_last_error(void);
uint32_t    abnn_abi_version(void);
/* Fill *p with the reference's constants under one of the two profiles. */
int abnn_default_params(abnn_params* p, uint32_t profile);

/* ---- lifetime: Brain::Brain + build_pipeline + build_buffers (brain.cpp:21-69) ------------- */
int  abnn_create(const abnn_params* p, abnn_handle** out);
void abnn_destroy(abnn_handle* h);                               /* Brain::release_all (brain.cpp:29-34) */
int  abnn_get_info(abnn_handle* h, abnn_info* out);              /* getters brain.h:48-52      */

/* ---- multi-GPU plumbing (new; the reference is single-device). One NCCL communicator per handle.
 * id is the 128-byte ncclUniqueId made by rank 0 and distributed by the host (any transport).
 * GPU-count invariance: the partition (abnn_partition), abnn_upload_synapses of one global table and the snapshot rule
 * for lastFired[src] do not depend on world_size; the RESULTS of a run do — rank is part of the Philox counter, ticks
 * are clock + i*world_size + rank, and ABNN_GRAPH_ER_BETA draws each rank's edges over its own neuron slice. A W-rank
 * run is reproducible for a given (seed, W) and is checked against the W-shard oracle.            */
int abnn_comm_unique_id(void* id128);
int abnn_comm_init(abnn_handle* h, const void* id128);

/* ---- graph: build_random_graph (brain-engine.cpp:31-53) / Brain::load|save (brain.cpp:161-178) */
int abnn_init_graph(abnn_handle* h, uint32_t kind, uint64_t seed);
/* Host table of the WHOLE graph (n records); the rank keeps, in order, the records whose dst it owns.
 * ABNN_ERR_INVALID if a kept record names a source neuron >= n_neuron, or (single rank) a destination >= n_neuron: the
 * kernels index the per-neuron arrays with both. The handle is then left with an EMPTY table. Same for abnn_load_bnn. */
int abnn_upload_synapses(abnn_handle* h, const abnn_synapse* syn, uint64_t n);
/* This rank's live records, in table order. *n_out = count; fails with ABNN_ERR_CAPACITY if cap is short. */
int abnn_download_synapses(abnn_handle* h, abnn_synapse* out, uint64_t cap, uint64_t* n_out);
/* .bnn v1: u32 N_SYN, u32 N_NRN, N_SYN x 16-byte records, no padding (brain.cpp:161-167). save writes the LIVE record
 * count; load wants N_NRN equal to the handle's and N_SYN <= its capacity (ABNN_ERR_SHAPE otherwise, where the reference
 * throws: brain.cpp:174) and the whole table present in the file before it touches the device table. */
int abnn_save_bnn(abnn_handle* h, const char* path);
int abnn_load_bnn(abnn_handle* h, const char* path);

/* .bnn v2 (new; README.md:237-241 lists what v1 lacks): everything needed to resume EXACTLY where the run
 * stopped — this rank's records, lastFired / lastVisited / snapshot, clock, pass and event counters,
 * reward, r-bar, read-out filter state, staged growth candidates. Header "BNN2", shape-checked like v1.
 * One file per rank. Both synchronise. load refuses (ABNN_ERR_SHAPE) a file written under other semantics — sampler,
 * release_rng, clock_mode, src_view, rbar_mode, sample_block, table_order, seed, window_pre, refractory, teacher_gap,
 * budget, track_visits, shape — while exec_mode, learning rates and placement may differ; the file is checked for
 * completeness before any device state is overwritten. */
int abnn_save_state(abnn_handle* h, const char* path);
int abnn_load_state(abnn_handle* h, const char* path);

/* ---- per-pass operations ---------------------------------------------------------------- */
/* Brain::inject_inputs (brain.cpp:73-83): input i spikes at `now` iff u < hz*kTickNS*NSEC_PER_SEC*v[i]. */
int abnn_inject_inputs(abnn_handle* h, const float* v, uint32_t n, float hz);
/* Teacher forcing (brain-engine.cpp:119-134): output o spikes iff u < expected[o]*rate and it is
 * not within teacher_gap of its last spike. */
int abnn_teacher_force(abnn_handle* h, const float* expected, uint32_t n, float rate);
/* Raw write through reward_buffer() (brain-engine.cpp:180-182). */
int abnn_set_reward(abnn_handle* h, float reward);
int abnn_get_reward(abnn_handle* h, float* reward, float* rbar);         /* synchronises */
/* Brain::encode_traversal + commit (brain.cpp:87-122, brain-engine.cpp:136-141): one pass of
 * `events` events. Asynchronous when stats == NULL; otherwise synchronises and fills *stats. */
int abnn_run_pass(abnn_handle* h, uint64_t events, abnn_pass_stats* stats);
int abnn_sync(abnn_handle* h);                                           /* waitUntilCompleted */
/* One whole engine pass without a host round trip — BrainEngine::run_one_pass (brain-engine.cpp:108-190):
 * stage the stimulus frame, inject_inputs(in, hz), teacher forcing (expected, teacher_rate), `events`
 * traversal events, timestamp exchange, read-out step with loss/reward. Equivalent to abnn_inject_inputs +
 * abnn_teacher_force + abnn_run_pass(NULL) + abnn_readout_step; the device work is recorded into a CUDA graph once a
 * call repeats the state of the one before it (same events, table sizes and word state: the third call of a steady
 * run) and replayed while that stays the same: PARALLEL and EXACT execution on single-GPU
 * handles, PARALLEL on sharded handles with exchange = ABNN_EXCHANGE_PEER (with the NCCL exchange sharded handles
 * enqueue the same sequence eagerly).
 * Asynchronous when rates == NULL; otherwise writes the n_output filtered rates and synchronises. */
int abnn_engine_step(abnn_handle* h, const float* in, const float* expected, float hz, float teacher_rate,
                     uint64_t events, float* rates);
/* Device-side stopwatch on the handle's stream (CUDA events; the stream is private to the handle, so
 * callers cannot time it with their own events): mark slot 0..7, then read the time between two marks.
 * abnn_timer_elapsed synchronises. */
int abnn_timer_mark(abnn_handle* h, uint32_t slot);
int abnn_timer_elapsed(abnn_handle* h, uint32_t slot_from, uint32_t slot_to, double* ms);
/* Brain::read_outputs (brain.cpp:145-157): spikes[o] = output o fired during the last pass. Synchronises. */
int abnn_read_outputs(abnn_handle* h, uint8_t* spikes, uint32_t n);
/* Rate EMA + RateFilter::process + peak normalise (+ loss/reward every reward_window passes when
 * expected != NULL): brain-engine.cpp:145-186. Writes the normalised smooth rates. Synchronises. */
int abnn_readout_filtered(abnn_handle* h, const float* expected, float* rates, uint32_t n);
/* Same read-out step, left on the device (no host copy, no synchronisation). */
int abnn_readout_step(abnn_handle* h, const float* expected, uint32_t n);
int abnn_get_loss(abnn_handle* h, double* last_loss, uint64_t* windows_done);   /* synchronises */

/* ---- structural plasticity (README.md:120-127): stable prune-compaction, then ordered append. Collective on sharded
 * handles (growth candidates are allgathered): a growth-buffer overflow on ANY rank makes every rank return
 * ABNN_ERR_CAPACITY together (the candidates of that interval are dropped). Synchronises. */
int abnn_prune_and_grow(abnn_handle* h, abnn_structural_stats* stats);

/* ---- raw state access: last_fired_buffer()/clock_buffer() (brain.h:54-58). All synchronise. - */
int abnn_download_timestamps(abnn_handle* h, uint64_t* last_fired, uint64_t* last_visited); /* n_neuron each; NULL to skip */
int abnn_upload_timestamps(abnn_handle* h, const uint64_t* last_fired, const uint64_t* last_visited);
/* Inspection: the 32-bit pre-spike gate words the line kernel will read in the NEXT pass (n_neuron values;
 * word = clamp(window_pre - (clock - lastFired_snapshot) + 1, 0, 2^32-2), traversal.cu:k_build_slack).
 * *valid_out = 0 when they have not been prepared yet (they are rebuilt at the start of the pass: first
 * pass, after abnn_set_clock / abnn_upload_timestamps, single-GPU handles, non-PARALLEL modes). Synchronises. */
int abnn_download_gate_words(abnn_handle* h, uint32_t* words, uint32_t* valid_out);
int abnn_get_clock(abnn_handle* h, uint64_t* clock);
int abnn_set_clock(abnn_handle* h, uint64_t clock);

/* ---- host-only helpers (no device needed) -------------------------------------------------- */
/* Flat-key YAML manifest -> abnn_params (abnn/manifests/simple.yml:3-12; the reference bundles the file but
 * never parses these keys). *p must already hold defaults (abnn_default_params). Top-level `key: value`
 * lines only; nested blocks, lists and unknown keys are skipped; numbers may carry `_` separators
 * (`20_000` — a string under YAML 1.2, which is why the reference's own parser would not read it).
 * Reference keys: neurons (total) -> n_hidden = neurons - n_input - n_output, synapses -> n_syn,
 * tau_LTP -> window_pre, alpha_LTP -> a_ltp, alpha_LTD -> a_ltd, w_min, w_max, rng_seed -> seed;
 * tau_LTD and steps are accepted and returned through the optional outputs (no kernel parameter uses
 * them, in the reference or here). Any abnn_params field name is accepted as a key too. */
int abnn_params_from_manifest(const char* path, abnn_params* p, uint64_t* steps_out, uint64_t* tau_ltd_out);
/* Destination-neuron range of `rank`: [lo, hi), slices of ceil(n_neuron/world). */
int abnn_partition(uint64_t n_neuron, uint32_t world_size, uint32_t rank, uint64_t* lo, uint64_t* hi);
/* Events rank r executes in a pass of `events` when shard r holds n_local of n_global synapses and
 * `before` synapses live on lower ranks: [first, first+count). */
int abnn_event_share(uint64_t events, uint64_t n_global, uint64_t before, uint64_t n_local,
                     uint64_t* first, uint64_t* count);
/* Philox4x32-10 (Salmon et al. 2011), the sampler of every random draw in this library. */
void abnn_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* ABNN_H_ */
