// include/abnn_brain.hpp — host C++17 façade over the C-ABI (include/abnn.h), header-only.
//
// Mirrors the reference's engine surface for the traversal hot path — same class names, method names,
// argument meaning and error behaviour — so that the reference's call sites compile against it:
//   Brain            abnn/src/core/brain/brain.h:24-83      (device-state owner; was Metal buffers + pipelines)
//   BrainEngine      abnn/src/core/brain-engine.h:33-85     (per-pass loop, async worker, model load/save)
//   StimulusProvider abnn/src/stimulus/stimulus-provider.h:20-33
//   FunctionalDataset abnn/src/stimulus/functional-dataset.{h,cpp}
// Differences, all forced by the boundary: no MTL::Device / MTL::CommandBuffer arguments (the handle owns
// its CUDA stream), counts are 64-bit, hyper-parameters come from abnn_params instead of #defines, and
// run_one_pass() is public. Link with -labnn_b200. Nothing here computes the hot path on the host.
#pragma once
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "abnn.h"

namespace abnn_b200 {

// constants.h:2-19 of the reference
constexpr uint32_t NUM_INPUTS = 256, NUM_OUTPUTS = 256;
constexpr uint64_t NUM_HIDDEN = 5000000ull, NUM_SYN = 1000000000ull, EVENTS_PER_PASS = 150000000ull;
constexpr float    INPUT_RATE_HZ = 1000.0f;
constexpr double   INPUT_SIN_WAVE_FREQUENCY = 0.5, dT_SEC = 0.0009;

using SynapsePacked = abnn_synapse;                     // brain.h:21

struct Error : std::runtime_error {
    int status;
    Error(int s, const char* where) : std::runtime_error(std::string(where) + ": " + abnn_last_error()), status(s) {}
};
inline void check(int status, const char* where) { if (status != ABNN_OK) throw Error(status, where); }

inline abnn_params default_params(uint32_t profile = ABNN_PROFILE_NORTH_STAR)
{
    abnn_params p;
    check(abnn_default_params(&p, profile), "abnn_default_params");
    return p;
}

class Brain {
public:
    // Brain::Brain (brain.cpp:21-27). `params` (optional) carries every other knob; shape fields are overwritten.
    Brain(uint32_t nInput, uint32_t nOutput, uint64_t nHidden, uint64_t nSynapses, uint64_t eventsPerPass,
          const abnn_params* params = nullptr)
        : p_(params ? *params : default_params()), EVENTS_(eventsPerPass)
    {
        p_.n_input = nInput; p_.n_output = nOutput; p_.n_hidden = nHidden; p_.n_syn = nSynapses;
    }
    ~Brain() { if (h_) abnn_destroy(h_); }              // release_all (brain.cpp:29-34)
    Brain(const Brain&) = delete;
    Brain& operator=(const Brain&) = delete;

    /* one-time initialisation */
    void build_pipeline() {}                            // kernels are compiled into libabnn_b200.so (brain.cpp:38-48)
    void build_buffers() { if (!h_) check(abnn_create(&p_, &h_), "abnn_create"); }       // brain.cpp:52-69

    /* per-pass operations */
    void encode_traversal(abnn_pass_stats* stats = nullptr)                              // brain.cpp:87-122 (+ commit)
    { check(abnn_run_pass(need(), EVENTS_, stats), "abnn_run_pass"); }
    void wait_until_completed() { check(abnn_sync(need()), "abnn_sync"); }               // brain-engine.cpp:141
    void inject_inputs(const std::vector<float>& vals, float hz)                         // brain.cpp:73-83
    { check(abnn_inject_inputs(need(), vals.data(), (uint32_t)vals.size(), hz), "abnn_inject_inputs"); }
    void teacher_force(const std::vector<float>& expected, float rate)                   // brain-engine.cpp:119-134
    { check(abnn_teacher_force(need(), expected.data(), (uint32_t)expected.size(), rate), "abnn_teacher_force"); }
    std::vector<bool> read_outputs()                                                     // brain.cpp:145-157
    {
        std::vector<uint8_t> s(p_.n_output);
        check(abnn_read_outputs(need(), s.data(), p_.n_output), "abnn_read_outputs");
        return std::vector<bool>(s.begin(), s.end());
    }
    // rate EMA + RateFilter::process + peak normalise (+ windowed loss -> reward): brain-engine.cpp:145-186
    std::vector<float> readout_filtered(const std::vector<float>* expected)
    {
        std::vector<float> r(p_.n_output);
        check(abnn_readout_filtered(need(), expected ? expected->data() : nullptr, r.data(), p_.n_output), "abnn_readout_filtered");
        return r;
    }
    // inject + teacher forcing + pass + read-out step in one enqueue (abnn_engine_step; CUDA-graph replay)
    std::vector<float> engine_step(const std::vector<float>& in, const std::vector<float>& expected, float hz, float teacherRate)
    {
        std::vector<float> r(p_.n_output);
        check(abnn_engine_step(need(), in.data(), expected.data(), hz, teacherRate, EVENTS_, r.data()), "abnn_engine_step");
        return r;
    }
    void set_reward(float r) { check(abnn_set_reward(need(), r), "abnn_set_reward"); }   // reward_buffer() poke, brain-engine.cpp:180-182

    /* graph */
    void build_random_graph(uint64_t seed = 1) { check(abnn_init_graph(need(), ABNN_GRAPH_REFERENCE, seed), "abnn_init_graph"); }   // brain-engine.cpp:31-53
    void init_graph(uint32_t kind, uint64_t seed) { check(abnn_init_graph(need(), kind, seed), "abnn_init_graph"); }
    void upload_synapses(const std::vector<SynapsePacked>& s) { check(abnn_upload_synapses(need(), s.data(), s.size()), "abnn_upload_synapses"); }
    std::vector<SynapsePacked> download_synapses()
    {
        abnn_info i; check(abnn_get_info(need(), &i), "abnn_get_info");
        std::vector<SynapsePacked> s(i.n_syn_local);
        uint64_t n = 0;
        check(abnn_download_synapses(h_, s.data(), s.size(), &n), "abnn_download_synapses");
        return s;
    }

    /* persistence: .bnn v1 = u32 N_SYN, u32 N_NRN, records (brain.cpp:161-178) */
    void save(std::ostream& os)
    {
        const std::vector<SynapsePacked> s = download_synapses();
        const uint32_t hdr[2] = {(uint32_t)s.size(), (uint32_t)n_neuron()};
        os.write(reinterpret_cast<const char*>(hdr), sizeof hdr);
        os.write(reinterpret_cast<const char*>(s.data()), (std::streamsize)(s.size() * sizeof(SynapsePacked)));
    }
    void load(std::istream& is)
    {
        uint32_t hdr[2] = {0, 0};
        is.read(reinterpret_cast<char*>(hdr), sizeof hdr);
        if (!is || hdr[0] != p_.n_syn || hdr[1] != n_neuron()) throw Error(ABNN_ERR_SHAPE, "Brain::load");   // brain.cpp:174
        std::vector<SynapsePacked> s(hdr[0]);
        is.read(reinterpret_cast<char*>(s.data()), (std::streamsize)(s.size() * sizeof(SynapsePacked)));
        if (!is) throw Error(ABNN_ERR_IO, "Brain::load");
        upload_synapses(s);
    }
    void save(const std::string& path) { check(abnn_save_bnn(need(), path.c_str()), "abnn_save_bnn"); }
    void load(const std::string& path) { check(abnn_load_bnn(need(), path.c_str()), "abnn_load_bnn"); }

    /* getters (brain.h:48-52) */
    uint32_t n_input() const { return p_.n_input; }
    uint32_t n_output() const { return p_.n_output; }
    uint64_t n_hidden() const { return p_.n_hidden; }
    uint64_t n_neuron() const { return (uint64_t)p_.n_input + p_.n_output + p_.n_hidden; }
    uint64_t n_syn() { abnn_info i; check(abnn_get_info(need(), &i), "abnn_get_info"); return i.n_syn_global; }
    uint64_t clock() { uint64_t c = 0; check(abnn_get_clock(need(), &c), "abnn_get_clock"); return c; }   // clock_buffer()
    abnn_handle* handle() { return need(); }
    const abnn_params& params() const { return p_; }
    abnn_params& params() { return p_; }                // editable until build_buffers()

private:
    abnn_handle* need() { if (!h_) throw Error(ABNN_ERR_INVALID, "Brain: build_buffers() has not been called"); return h_; }
    abnn_params p_;
    uint64_t EVENTS_;
    abnn_handle* h_ = nullptr;
};

class StimulusProvider {                                // stimulus-provider.h:20-33
public:
    virtual ~StimulusProvider() = default;
    virtual std::vector<float> nextInput() = 0;
    virtual std::vector<float> nextExpected() = 0;
    virtual double time() const = 0;
};

class FunctionalDataset : public StimulusProvider {     // functional-dataset.cpp:9-52
public:
    FunctionalDataset(uint32_t nInput, uint32_t nOutput, double dtSec, double freqHz,
                      std::function<float(float)> funcInput, std::function<float(float)> funcExpected)
        : nInput_(nInput), nOutput_(nOutput), dt_(dtSec), tSec_(0.0), fHz_(freqHz), phase_(0.0),
          funcInput_(std::move(funcInput)), funcExpected_(std::move(funcExpected)) {}
    std::vector<float> nextInput() override
    {
        phase_ += fHz_ * dt_;
        if (phase_ > 1.0) phase_ -= 1.0;
        tSec_ += dt_;
        std::vector<float> v(nInput_);
        for (uint32_t i = 0; i < nInput_; ++i) {
            const double x = double(i) / nInput_;
            v[i] = funcInput_(float(2.0 * M_PI * (x + phase_)));
        }
        return v;
    }
    std::vector<float> nextExpected() override
    {
        std::vector<float> v(nOutput_);
        for (uint32_t i = 0; i < nOutput_; ++i) {
            const double x = double(i) / nOutput_;
            const double s = funcExpected_(float(2.0 * M_PI * (x + phase_)));
            v[i] = float(s);
        }
        return v;
    }
    double time() const override { return tSec_; }

private:
    uint32_t nInput_, nOutput_;
    double dt_, tSec_, fHz_, phase_;
    std::function<float(float)> funcInput_, funcExpected_;
};

// Logger (abnn/src/core/singletons/logger.{h,cpp}): Octave/MATLAB animation script of input vs filtered
// output (one frame per log_samples call) and the loss EMA (beta = 0.98), truncated every 10 losses.
class Logger {
public:
    Logger(int nInput, int nOutput, const std::string& file = "abnn_session.m") : nIn_(nInput), nOut_(nOutput), file_(file) { open(); }
    void log_samples(const std::vector<float>& in, const std::vector<float>& out)          // logger.cpp:25-56
    {
        if (!mat_) return;
        mat_ << "clf;\nhold on;\nylim([-1 1]);\n";
        mat_ << "xo = [ "; for (int i = 0; i < nOut_; ++i) mat_ << i << " "; mat_ << "];\n";
        mat_ << "x = [ ";  for (size_t i = 0; i < in.size(); ++i) mat_ << i << " "; mat_ << "];\n";
        mat_ << "y = [ ";  for (size_t i = 0; i < in.size(); ++i) mat_ << in[i] << " "; mat_ << "];\n";
        mat_ << "\nz=[";   for (size_t i = 0; i < out.size(); ++i) { if (i) mat_ << ","; mat_ << out[i]; }
        mat_ << "];title('Output');\n";
        mat_ << "scatter(x,y,[],[],[0,0,1]);\nscatter(xo,z,[],[],[0,1,0]);\nhold off; pause(0.03);\n\n";
        mat_.flush();
        ++frames_;
    }
    void accumulate_loss(double loss)                                                       // logger.cpp:59-69
    {
        ema_ = step_ == 0 ? loss : beta_ * ema_ + (1.0 - beta_) * loss;
        ++step_;
        if (verbose) std::printf("EMA-Loss: %g  Raw loss: %g\n", ema_, loss);
        if (step_ % 10 == 0) flush();
    }
    void flush() { mat_.flush(); mat_.close(); open(); }                                   // logger.cpp:71-84
    double ema() const { return ema_; }
    uint64_t losses() const { return step_; }
    uint64_t frames() const { return frames_; }
    bool verbose = false;

private:
    void open() { mat_.open(file_, std::ios::trunc); if (mat_) mat_ << "% ABNN animated session\n"; }
    int nIn_, nOut_;
    std::string file_;
    std::ofstream mat_;
    double ema_ = 0.0, beta_ = 0.98;
    uint64_t step_ = 0, frames_ = 0;
};

class BrainEngine {                                     // brain-engine.h:33-85
public:
    BrainEngine(uint32_t nInput, uint32_t nOutput, uint64_t eventsPerPass = EVENTS_PER_PASS,
                const abnn_params* params = nullptr, const std::string& modelFile = "model.bnn")
        : nIn_(nInput), nOut_(nOutput), eventsPerPass_(eventsPerPass), modelFile_(modelFile)
    {
        abnn_params p = params ? *params : default_params();
        brain_ = std::make_unique<Brain>(nIn_, nOut_, p.n_hidden, p.n_syn, eventsPerPass_, &p);   // brain-engine.cpp:66
        brain_->build_pipeline();
        brain_->build_buffers();
        rewardWindow_ = p.reward_window;
        if (!load_model()) {                            // brain-engine.cpp:72-75
            brain_->build_random_graph(1);
            save_model();
        }
    }
    ~BrainEngine() { stop_async(); }

    void set_stimulus(std::shared_ptr<StimulusProvider> s) { stim_ = std::move(s); }      // brain-engine.cpp:105

    void start_async()                                  // brain-engine.cpp:193-202
    {
        if (running_.exchange(true)) return;
        worker_ = std::thread([this] { while (running_) run_one_pass(); });
    }
    void stop_async()                                   // brain-engine.cpp:203-207
    {
        if (!running_.exchange(false)) return;
        if (worker_.joinable()) worker_.join();
    }

    bool load_model(const std::string& nm = "")         // brain-engine.cpp:85-97
    {
        const std::string f = nm.empty() ? modelFile_ : nm;
        if (f.empty()) return false;
        std::ifstream is(f, std::ios::binary);
        if (!is) return false;
        try { brain_->load(is); } catch (const Error&) { return false; }
        return true;
    }
    bool save_model(const std::string& nm = "")         // brain-engine.cpp:99-102
    {
        const std::string f = nm.empty() ? modelFile_ : nm;
        if (f.empty()) return false;
        std::ofstream os(f, std::ios::binary);
        if (!os) return false;
        brain_->save(os);
        return bool(os);
    }

    // One synchronous pass, the order of operations of brain-engine.cpp:108-190. Returns the normalised,
    // filtered output rates (what the reference hands to its logger).
    std::vector<float> run_one_pass()
    {
        const std::vector<float> in = stim_->nextInput();                                 // :114
        const std::vector<float> expected = stim_->nextExpected();                        // :115
        // :117 inject_inputs, :119-134 teacher forcing on alternate passes (`static bool even`), :136-141 traversal,
        // :143-186 read-out + loss/reward — one enqueue, one synchronisation
        std::vector<float> smooth = brain_->engine_step(in, expected, INPUT_RATE_HZ, even_ ? 1.0f : 0.0f);
        even_ = !even_;
        ++step_;
        if (logger_) {
            if (step_ % 100 == 0) logger_->log_samples(in, smooth);                       // :166-168
            if (rewardWindow_ && step_ % rewardWindow_ == 0) {                            // :173-186 (loss computed on the device)
                double loss = 0.0; uint64_t windows = 0;
                check(abnn_get_loss(brain_->handle(), &loss, &windows), "abnn_get_loss");
                logger_->accumulate_loss(loss);
            }
        }
        return smooth;
    }
    // Attach the reference's Logger (writes abnn_session.m in the working directory by default).
    void enable_logger(const std::string& file = "abnn_session.m") { logger_ = std::make_unique<Logger>((int)nIn_, (int)nOut_, file); }
    Logger* logger() { return logger_.get(); }

    Brain& brain() { return *brain_; }
    uint64_t step() const { return step_; }

private:
    std::unique_ptr<Brain> brain_;
    std::unique_ptr<Logger> logger_;
    std::shared_ptr<StimulusProvider> stim_;
    uint32_t rewardWindow_ = 0;
    std::thread worker_;
    std::atomic<bool> running_{false};
    uint32_t nIn_, nOut_;
    uint64_t eventsPerPass_;
    std::string modelFile_;
    bool even_ = false;
    uint64_t step_ = 0;
};

}  // namespace abnn_b200
