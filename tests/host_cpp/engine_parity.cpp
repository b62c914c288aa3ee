// tests/host_cpp/engine_parity.cpp — drives the hot path from host C++17 exactly as the reference's
// ViewDelegate / BrainEngine do (view-delegate.cpp:30-44, brain-engine.cpp:108-190), through
// include/abnn_brain.hpp -> C-ABI -> CUDA. Writes every pass's filtered read-out and the final synapse
// table to a file; tests/test_host_cpp.py compares the bytes with the oracle.
//   engine_parity <out.bin> <passes> [model.bnn]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "abnn_brain.hpp"

using namespace abnn_b200;

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: engine_parity <out.bin> <passes> [model.bnn]\n"); return 2; }
    const int passes = std::atoi(argv[2]);
    try {
        abnn_params p = default_params(ABNN_PROFILE_NORTH_STAR);
        p.n_hidden = 2000; p.n_syn = 60000;
        p.exec_mode = ABNN_EXEC_SERIAL; p.clock_mode = ABNN_CLOCK_PER_PASS;
        p.window_pre = 5; p.refractory = 2; p.max_spikes_per_pass = 2560; p.reward_window = 10;
        BrainEngine engine(64, 64, 60000, &p, argc > 3 ? argv[3] : "");
        auto stim = std::make_shared<FunctionalDataset>(64, 64, dT_SEC, INPUT_SIN_WAVE_FREQUENCY,
                                                        [](float x) { return cos(x) * cos(x); },           // view-delegate.cpp:37-39
                                                        [](float x) { return 0.5f * sin(x) + 0.5f; });     // view-delegate.cpp:40-42
        engine.set_stimulus(stim);
        engine.enable_logger(std::string(argv[1]) + ".session.m");
        std::ofstream out(argv[1], std::ios::binary);
        for (int i = 0; i < passes; ++i) {
            const std::vector<float> r = engine.run_one_pass();
            out.write(reinterpret_cast<const char*>(r.data()), (std::streamsize)(r.size() * sizeof(float)));
        }
        const std::vector<SynapsePacked> table = engine.brain().download_synapses();
        out.write(reinterpret_cast<const char*>(table.data()), (std::streamsize)(table.size() * sizeof(SynapsePacked)));
        // stream save/load round trip (Brain::save / Brain::load, brain.cpp:161-178) and the shape check
        std::stringstream ss;
        engine.brain().save(ss);
        engine.brain().load(ss);
        if (engine.brain().download_synapses().size() != table.size()) { std::fprintf(stderr, "round trip changed the table\n"); return 1; }
        abnn_params q = p; q.n_hidden = 2001;
        Brain other(64, 64, q.n_hidden, q.n_syn, 60000, &q);
        other.build_buffers();
        std::stringstream s2;
        engine.brain().save(s2);
        bool threw = false;
        try { other.load(s2); } catch (const Error& e) { threw = e.status == ABNN_ERR_SHAPE; }
        if (!threw) { std::fprintf(stderr, "shape mismatch was not reported\n"); return 1; }
        // async worker (start_async / stop_async, brain-engine.cpp:193-207)
        const uint64_t before = engine.step();
        engine.start_async();
        while (engine.step() < before + 3) std::this_thread::yield();
        engine.stop_async();
        std::printf("ok passes=%d async_passes=%llu clock=%llu logger_losses=%llu ema=%.9g\n", passes,
                    (unsigned long long)(engine.step() - before), (unsigned long long)engine.brain().clock(),
                    (unsigned long long)engine.logger()->losses(), engine.logger()->ema());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "engine_parity: %s\n", e.what());
        return 1;
    }
    return 0;
}
