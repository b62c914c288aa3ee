// oracle/ref_pieces.cpp — TEST INFRASTRUCTURE ONLY.
// C wrappers around reference sources compiled VERBATIM from /root/reference (found through -I,
// never copied): core/output-filter/rate-filter.h (RateFilter), stimulus/functional-dataset.cpp
// (FunctionalDataset) and core/singletons/logger.cpp (Logger), #included below as translation-unit members. The two stimulus lambdas are
// the ones the app installs at abnn/src/view-delegate.cpp:37-42, restated character for character.
// Used to validate the restatements in oracle_b.cpp and to generate tests/golden/*.json
// (tests/golden/make_golden.py). Output: oracle/_ref/libref_pieces.so (git-ignored).
#include <cmath>
#include <cstdint>
#include <functional>
#include <memory>
#include <vector>
#include <iomanip>             // logger.cpp:65 uses std::setprecision without including it
#include "rate-filter.h"
#include "functional-dataset.h"
#include "functional-dataset.cpp"
#include "singletons/logger.cpp"

extern "C" {
void* refp_dataset_create(uint32_t nIn, uint32_t nOut, double dt, double f)
{
    return new FunctionalDataset(nIn, nOut, dt, f,
                                 [](float x){
                                     return cos(x) * cos(x);
                                 },
                                 [](float x){
                                     return 0.5f * sin(x) + 0.5f;
                                 });
}
void refp_dataset_destroy(void* d) { delete static_cast<FunctionalDataset*>(d); }
void refp_dataset_next_input(void* d, float* out)
{
    auto v = static_cast<FunctionalDataset*>(d)->nextInput();
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
}
void refp_dataset_next_expected(void* d, float* out)
{
    auto v = static_cast<FunctionalDataset*>(d)->nextExpected();
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
}
double refp_dataset_time(void* d) { return static_cast<FunctionalDataset*>(d)->time(); }

void* refp_filter_create(double tau, int useFIR, uint64_t firSize) { return new RateFilter(tau, useFIR != 0, firSize); }
void  refp_filter_destroy(void* f) { delete static_cast<RateFilter*>(f); }
void  refp_filter_process(void* f, const float* raw, uint32_t n, double dt, float* out)
{
    std::vector<float> r(raw, raw + n);
    auto v = static_cast<RateFilter*>(f)->process(r, dt);
    for (uint32_t i = 0; i < n; ++i) out[i] = v[i];
}

// Logger (logger.cpp:13-84): writes abnn_session.m into the CURRENT directory (the caller chdir()s first).
void* refp_logger_create(int nIn, int nOut) { return new Logger(nIn, nOut); }
void  refp_logger_destroy(void* l) { delete static_cast<Logger*>(l); }
void  refp_logger_log_samples(void* l, const float* in, uint32_t nIn, const float* out, uint32_t nOut)
{
    static_cast<Logger*>(l)->log_samples(std::vector<float>(in, in + nIn), std::vector<float>(out, out + nOut));
}
void  refp_logger_accumulate_loss(void* l, double loss) { static_cast<Logger*>(l)->accumulate_loss(loss); }
}
