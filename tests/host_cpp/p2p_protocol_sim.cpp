// tests/host_cpp/p2p_protocol_sim.cpp — host model of the flag protocol of the peer-memory exchange
// (abnn_b200/csrc/exchange.cu: k_p2p_done / k_p2p_push / k_p2p_wait), ranks as threads with random delays.
// It checks the two properties the protocol exists for, over many passes:
//   1. while a rank "traverses" pass k it only ever reads gate words of pass k (no peer overwrites them early);
//   2. when a rank starts pass k + 1 every word of its array is a pass-(k + 1) word (nothing arrives late);
// and that nobody deadlocks. It models the LOGIC (epochs, which flag gates what); the CUDA memory-ordering side
// (__threadfence_system, volatile polls) is not modelled — std::atomic with seq_cst stands in for it.
// Built and run by tests/test_host_cpp.py::test_p2p_protocol_model (g++, no GPU).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

typedef unsigned long long u64;
static const int W = 4, SLICE = 64, N = W * SLICE, PASSES = 400;
enum { DONE = 0, PUSHED = 8, EPOCH = 16, WORDS = 32 };

struct Rank {
    std::vector<std::atomic<unsigned>> slack;      // gate words of every neuron; value = pass the word was built for
    std::atomic<u64> flags[WORDS];
    Rank() : slack(N) { for (auto& f : flags) f = 0; for (auto& s : slack) s = 0; }
};
static Rank ranks[W];
static std::atomic<int> failures{0};

static void nap(std::mt19937& g, int max_us) { if (max_us) std::this_thread::sleep_for(std::chrono::microseconds(g() % max_us)); }

static void wait_all(Rank& me, int base, u64 epoch)
{
    for (int r = 0; r < W; ++r)
        while (me.flags[base + r].load() < epoch && !failures.load()) std::this_thread::yield();
}

static void run_rank(int me_id)
{
    std::mt19937 g(1234 + me_id);
    Rank& me = ranks[me_id];
    for (unsigned pass = 0; pass < (unsigned)PASSES; ++pass) {
        // ---- traversal of `pass`: reads random gate words, each must belong to this pass
        if (failures.load()) return;                                     // another rank failed: do not wait for it
        const int reads = 20 + g() % 200;
        for (int i = 0; i < reads; ++i) {
            const unsigned w = me.slack[g() % N].load();
            if (w != pass) { failures++; std::printf("rank %d pass %u read a word of pass %u\n", me_id, pass, w); return; }
            if ((g() & 63) == 0) nap(g, me_id == 1 ? 40 : 5);            // rank 1 is the slow one
        }
        // ---- k_p2p_done
        const u64 epoch = me.flags[EPOCH].load() + 1;
        me.flags[EPOCH].store(epoch);
        for (int r = 0; r < W; ++r) ranks[r].flags[DONE + me_id].store(epoch);
        // ---- k_p2p_push: wait for every DONE, store the owned slice everywhere, raise PUSHED
        wait_all(me, DONE, epoch);
        for (int n = me_id * SLICE; n < (me_id + 1) * SLICE; ++n)
            for (int r = 0; r < W; ++r) ranks[r].slack[n].store(pass + 1);
        nap(g, 3);
        for (int r = 0; r < W; ++r) ranks[r].flags[PUSHED + me_id].store(epoch);
        // ---- k_p2p_wait
        wait_all(me, PUSHED, epoch);
        for (int n = 0; n < N; ++n)
            if (me.slack[n].load() != pass + 1) { failures++; std::printf("rank %d starts pass %u with a stale word at %d\n", me_id, pass + 1, n); return; }
    }
}

int main()
{
    std::vector<std::thread> th;
    for (int r = 0; r < W; ++r) th.emplace_back(run_rank, r);
    for (auto& t : th) t.join();
    if (failures.load()) return 1;
    std::printf("ok %d ranks x %d passes\n", W, PASSES);
    return 0;
}
