// Port of the reference driver's decode step onto the AFF3CT-shaped wrapper.
//
// The reference loop ("main.cpp (alist)":411-441) is
//     m.decoder->decode_siho(b.LLRs, b.dec_bits);  ...  (*(m.decoder)).reset();
// with  H = tools::LDPC_matrix_handler::read(...)  and info_bits_pos taken from G (:333-340).
// This driver reads H (.alist / .qc), a text file of n_frames*N LLRs, the first information position,
// and prints the K decoded bits per frame -- enough to replay the reference's known-answer test
// ("main.cpp (alist)":443-462) through the same call sequence.
//
//   driver_siho <H file> <llr file> <info_first> <K> <n_ite> [rule: spa|ms|nms:<f>|oms:<o>] [layered|flooding] [Q: f32|i8|i16]
// With Q = i8 / i16 the LLR file holds integers and the decoder is Decoder_LDPC_BP<int, int8_t / int16_t>: the template
// parameter Q of Decoder_SISO_SIHO<B,Q> (BOOT/src/main.cpp:113) selects the fixed-point kernels; the kernel name goes to stderr.
// exit codes: 0 ok, 2 usage, 3 no CUDA device / decoder could not be created, 4 other error
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "qldpc_aff3ct.hpp"

using namespace qldpc;

template <typename Q>
static int decode_file(tools::Sparse_matrix H, const char *llr_file, int info_first, int K, int n_ite, const tools::Update_rule &rule,
                       bool layered)
{
    const int N = H.get_n_rows();
    std::vector<Q> llrs;
    std::ifstream in(llr_file);
    for (float v; in >> v;) llrs.push_back((Q)v);
    if (llrs.empty() || llrs.size() % (size_t)N != 0) throw tools::length_error("LLR file does not hold a multiple of N values");
    const int n_frames = (int)(llrs.size() / (size_t)N);
    std::vector<uint32_t> info_bits_pos(K);
    for (int i = 0; i < K; ++i) info_bits_pos[i] = (uint32_t)(info_first + i);
    module::Decoder_LDPC_BP<int, Q> decoder(K, N, n_ite, H, info_bits_pos, rule, /*enable_syndrome=*/true, /*syndrome_depth=*/1,
                                            n_frames, layered ? module::Schedule::horizontal_layered : module::Schedule::flooding);
    std::fprintf(stderr, "kernel=%s\n", decoder.kernel_name());
    std::vector<int> dec_bits((size_t)n_frames * K);
    decoder.decode_siho(llrs, dec_bits);
    decoder.reset();
    for (int f = 0; f < n_frames; ++f) {
        for (int i = 0; i < K; ++i) std::putchar('0' + dec_bits[(size_t)f * K + i]);
        std::printf(" iters=%d ok=%d\n", (int)decoder.last_iterations()[f], (int)decoder.last_syndrome_ok()[f]);
    }
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s <H> <llrs.txt> <info_first> <K> <n_ite> [rule] [layered|flooding] [f32|i8|i16]\n", argv[0]);
        return 2;
    }
    try {
        if (argc > 8 && std::strcmp(argv[8], "f32") != 0) {
            tools::Sparse_matrix Hq = tools::Sparse_matrix::read(argv[1]);
            tools::Update_rule r = tools::Update_rule_MS();
            const std::string rs = argv[6];
            if (rs.rfind("nms:", 0) == 0) r = tools::Update_rule_NMS(std::stof(rs.substr(4)));
            else if (rs.rfind("oms:", 0) == 0) r = tools::Update_rule_OMS(std::stof(rs.substr(4)));
            else if (rs == "spa") r = tools::Update_rule_SPA();
            const bool lay = std::strcmp(argv[7], "layered") == 0;
            if (std::strcmp(argv[8], "i8") == 0) return decode_file<int8_t>(Hq, argv[2], std::atoi(argv[3]), std::atoi(argv[4]), std::atoi(argv[5]), r, lay);
            if (std::strcmp(argv[8], "i16") == 0) return decode_file<int16_t>(Hq, argv[2], std::atoi(argv[3]), std::atoi(argv[4]), std::atoi(argv[5]), r, lay);
            return 2;
        }
        tools::Sparse_matrix H = tools::Sparse_matrix::read(argv[1]);
        const int N = H.get_n_rows();
        const int info_first = std::atoi(argv[3]), K = std::atoi(argv[4]), n_ite = std::atoi(argv[5]);
        tools::Update_rule rule = tools::Update_rule_SPA(H.get_cols_max_degree());
        if (argc > 6) {
            const std::string r = argv[6];
            if (r == "ms") rule = tools::Update_rule_MS();
            else if (r.rfind("nms:", 0) == 0) rule = tools::Update_rule_NMS(std::stof(r.substr(4)));
            else if (r.rfind("oms:", 0) == 0) rule = tools::Update_rule_OMS(std::stof(r.substr(4)));
        }
        const bool layered = argc > 7 && std::strcmp(argv[7], "layered") == 0;

        std::vector<float> llrs;
        std::ifstream in(argv[2]);
        for (float v; in >> v;) llrs.push_back(v);
        if (llrs.empty() || llrs.size() % (size_t)N != 0) throw tools::length_error("LLR file does not hold a multiple of N values");
        const int n_frames = (int)(llrs.size() / (size_t)N);

        std::vector<uint32_t> info_bits_pos(K);
        for (int i = 0; i < K; ++i) info_bits_pos[i] = (uint32_t)(info_first + i);

        module::Decoder_LDPC_BP<int, float> decoder(K, N, n_ite, H, info_bits_pos, rule, /*enable_syndrome=*/true,
                                                    /*syndrome_depth=*/1, n_frames,
                                                    layered ? module::Schedule::horizontal_layered : module::Schedule::flooding);
        std::vector<int> dec_bits((size_t)n_frames * K);
        decoder.decode_siho(llrs, dec_bits);
        decoder.reset();
        for (int f = 0; f < n_frames; ++f) {
            for (int i = 0; i < K; ++i) std::putchar('0' + dec_bits[(size_t)f * K + i]);
            std::printf(" iters=%d ok=%d\n", (int)decoder.last_iterations()[f], (int)decoder.last_syndrome_ok()[f]);
        }
        // wrong sizes raise, as in AFF3CT
        try {
            std::vector<int> too_short(3);
            decoder.decode_siho(llrs, too_short);
            return 4;
        } catch (const tools::length_error &) {
        }
        return 0;
    } catch (const tools::runtime_error &e) {
        std::fprintf(stderr, "runtime_error: %s\n", e.what());
        return 3;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 4;
    }
}
