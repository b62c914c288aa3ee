// tests/host_cpp/logger_parity.cpp — drives abnn_b200::Logger (include/abnn_brain.hpp) with a script written by
// tests/test_oracle.py so that abnn_session.m can be compared byte for byte with the reference's Logger compiled verbatim
// (oracle/ref_pieces.cpp). Usage: logger_parity <session file> <script.bin> <n_in> <n_out>. The script is a sequence of
// records: int32 op (0 = frame: n_in + n_out floats follow, 1 = loss: one double follows). After every record the
// session file is copied to <session file>.<record index>.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "abnn_brain.hpp"

int main(int argc, char** argv)
{
    if (argc < 5) return 2;
    const std::string path = argv[1];
    const int nIn = std::atoi(argv[3]), nOut = std::atoi(argv[4]);
    std::ifstream script(argv[2], std::ios::binary);
    if (!script) return 3;
    abnn_b200::Logger log(nIn, nOut, path);
    std::vector<float> in(nIn), out(nOut);
    int32_t op = 0;
    for (int k = 0; script.read(reinterpret_cast<char*>(&op), 4); ++k) {
        if (op == 0) {
            script.read(reinterpret_cast<char*>(in.data()), nIn * 4);
            script.read(reinterpret_cast<char*>(out.data()), nOut * 4);
            log.log_samples(in, out);
        } else {
            double loss = 0;
            script.read(reinterpret_cast<char*>(&loss), 8);
            log.accumulate_loss(loss);
        }
        std::ifstream src(path, std::ios::binary);
        std::ofstream dst(path + "." + std::to_string(k), std::ios::binary);
        dst << src.rdbuf();
    }
    std::printf("ema %.17g\n", log.ema());
    return 0;
}
