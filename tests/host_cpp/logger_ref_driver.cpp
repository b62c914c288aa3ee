// tests/host_cpp/logger_ref_driver.cpp — the same script as logger_parity.cpp, played to the REFERENCE's Logger compiled
// verbatim (oracle/_ref/libref_pieces.so: refp_logger_*). A separate process, because the reference writes
// abnn_session.m into its current directory. Usage: logger_ref_driver <script.bin> <n_in> <n_out>; after every record
// ./abnn_session.m is copied to ./ref.<record index>.
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

extern "C" {
void* refp_logger_create(int nIn, int nOut);
void refp_logger_destroy(void* l);
void refp_logger_log_samples(void* l, const float* in, uint32_t nIn, const float* out, uint32_t nOut);
void refp_logger_accumulate_loss(void* l, double loss);
}

int main(int argc, char** argv)
{
    if (argc < 4) return 2;
    const int nIn = std::atoi(argv[2]), nOut = std::atoi(argv[3]);
    std::ifstream script(argv[1], std::ios::binary);
    if (!script) return 3;
    void* log = refp_logger_create(nIn, nOut);
    std::vector<float> in(nIn), out(nOut);
    int32_t op = 0;
    for (int k = 0; script.read(reinterpret_cast<char*>(&op), 4); ++k) {
        if (op == 0) {
            script.read(reinterpret_cast<char*>(in.data()), nIn * 4);
            script.read(reinterpret_cast<char*>(out.data()), nOut * 4);
            refp_logger_log_samples(log, in.data(), nIn, out.data(), nOut);
        } else {
            double loss = 0;
            script.read(reinterpret_cast<char*>(&loss), 8);
            refp_logger_accumulate_loss(log, loss);
        }
        std::ifstream src("abnn_session.m", std::ios::binary);
        std::ofstream dst("ref." + std::to_string(k), std::ios::binary);
        dst << src.rdbuf();
    }
    refp_logger_destroy(log);
    return 0;
}
