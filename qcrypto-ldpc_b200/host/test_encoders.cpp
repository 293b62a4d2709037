// Replays the reference's encoding known-answer vector ("main.cpp (alist)":443-462: data[504] -> encoded[1008]) through the
// AFF3CT-shaped encoder classes of qldpc_aff3ct.hpp.
//   test_encoders <H.alist> <G.alist> <data.txt> <encoded.txt>      exit 0: both encoders reproduce `encoded`; 3: no device
#include <cstdio>
#include <fstream>

#include "qldpc_aff3ct.hpp"

using namespace qldpc;

static std::vector<int> read_bits(const char *path)
{
    std::ifstream in(path);
    std::vector<int> v;
    int b;
    while (in >> b) v.push_back(b);
    return v;
}

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    try {
        const std::vector<int> data = read_bits(argv[3]), want = read_bits(argv[4]);
        const int K = (int)data.size(), N = (int)want.size();
        auto H = tools::Sparse_matrix::read(argv[1]);
        module::Encoder_LDPC<int> from_g(K, N, argv[2], 2);
        module::Encoder_LDPC_from_H<int> from_h(K, N, H, 2);
        std::vector<int> u(2 * (size_t)K), x(2 * (size_t)N);
        for (int i = 0; i < K; ++i) { u[i] = data[i]; u[K + i] = 1 - data[i]; }
        int bad = 0;
        for (int which = 0; which < 2; ++which) {
            if (which == 0) from_g.encode(u, x); else from_h.encode(u, x);
            const auto pos = which == 0 ? from_g.get_info_bits_pos() : from_h.get_info_bits_pos();
            for (int i = 0; i < N; ++i) bad += x[i] != want[i];                       // frame 0: the reference's codeword
            for (int i = 0; i < K; ++i) bad += x[N + pos[i]] != u[K + i];             // frame 1: systematic at info_bits_pos
        }
        std::printf("mismatches %d\n", bad);
        return bad ? 1 : 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 3;
    }
}
