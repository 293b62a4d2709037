// The reference's QKD simulation driver for 5G quasi-cyclic codes, "main.cpp (5g-qc)" (VAR/main.cpp (5g-qc):365-545), on
// the GPU decoder: same command line, same QBER loop, same block puncturing to a target efficiency, same LLR construction,
// same report lines.  What changed is one line of the reference -- the decoder is module::Decoder_LDPC_BP (qldpc_aff3ct.hpp)
// instead of AFF3CT's -- and that the frames of one QBER step are decoded in ONE decode_siho call (AFF3CT's n_frames).
//
//   driver_5gqc <file.qc> <expansion_factor> <desired_reconciliation_efficiency> [frames_per_ber] [rule] [Q] [n_ite] [seed]
//     rule: spa (default, the reference's case 9) | ms | nms:<factor> | oms:<offset>, optionally ",layered"
//     Q   : f32 (default) | i8 | i16  -- integer Q: LLRs scaled by 4 and rounded, confirmed bits at 127 / 2047
// exit codes: 0 ok, 2 usage, 3 decoder error (no sm_100 device: there is no CPU fallback)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "qldpc_aff3ct.hpp"

using namespace qldpc;

#define CONFIRMED_BIT_LLR (-std::log(1e-10 / (1 - 1e-10)))                       // (5g-qc):33
// the reference's macros evaluate in float (ber is a float, std::log2(float) is float): the recorded puncture counts
// (92 at ber = 1e-10, where 1 - ber == 1.0f) are only reproduced in float
static float h(float q) { return (-q) * std::log2(q) - (1 - q) * std::log2(1 - q); }                    // :36
static float f_eff(float ratio, float q) { return (1 - ratio) / h(q); }                                // :38
static float k_using_f(float eff, float q, float fer) { return (1 - (1 + eff) * h(q)) * (1 - fer); }   // :40

template <typename Q>
static int run(const std::string &path, int expansion_factor, float desired_eff, int frames, const tools::Update_rule &rule,
               module::Schedule sched, int n_ite, unsigned seed, float q_scale, float q_confirmed)
{
    tools::Sparse_matrix H = tools::Sparse_matrix::read(path);                    // :389
    const int N = H.get_n_rows(), K = N - H.get_n_cols();                          // :375-376
    module::Encoder_LDPC_from_QC<int> encoder(K, N, H, frames);                    // :177
    const std::vector<uint32_t> info_bits_pos = encoder.get_info_bits_pos();       // :183
    module::Decoder_LDPC_BP<int, Q> decoder(K, N, n_ite, H, info_bits_pos, rule, true, 1, frames, sched);   // :242
    std::printf("# * Simulation parameters: \n#    ** Info. bits (K) = %d\n#    ** Frame size (N) = %d\n#    ** Expansion Factor = %d\n"
                "#    ** kernel = %s\n", K, N, expansion_factor, decoder.kernel_name());
    std::printf("#             QBER   ||     Iter | Bit Errs |Frame Err | BER      | FER      \n");
    std::mt19937 rng(seed);
    std::vector<int> ref_bits((size_t)K * frames), enc_bits((size_t)N * frames), dec_bits((size_t)K * frames);
    std::vector<Q> LLRs((size_t)N * frames);
    for (float ber = 0.00f; ber <= 0.11f; ber += 0.01f) {                          // :448, p.ber_min / max / step (:53-55)
        if (ber == 0) ber = 1e-10f;                                                // :450-453
        int bits_to_puncture = (int)(N - K - desired_eff * h(ber) * (float)K);     // :459
        bits_to_puncture = (bits_to_puncture / expansion_factor) * expansion_factor;   // :461
        const float R = (float)K / (float)(N - bits_to_puncture);
        const float ratio = 1 - (float)(N - K - bits_to_puncture) / (float)K;      // :464, R(PAR_BITS, INFO_BITS)
        const float efficiency = f_eff(ratio, ber), key_rate = k_using_f(efficiency, ber, 0.0f);
        std::printf("\nPunct. Bits (Round down): %d| Reconcil. Effic.: %g| Key Rate (at 0%% FER): %g| Final key len: %g| Code rate: %g\n",
                    bits_to_puncture, efficiency, key_rate, (float)K * key_rate, R);   // :478-483
        std::bernoulli_distribution bit(0.5), flip(ber);
        for (auto &b : ref_bits) b = bit(rng) ? 1 : 0;                             // m.source->generate (:502)
        encoder.encode(ref_bits, enc_bits);                                        // :503
        const float llr_noisy = std::log((1 - ber) / ber);                         // Modem_OOK_BSC::demodulate (:511)
        for (int f = 0; f < frames; ++f) {
            const int *x = enc_bits.data() + (size_t)f * N;
            Q *L = LLRs.data() + (size_t)f * N;
            for (int i = 0; i < N; ++i) {
                const int y = x[i] ^ (flip(rng) ? 1 : 0);                          // Channel_binary_symmetric (:506)
                float v = y ? -llr_noisy : llr_noisy;
                if (i >= K && i < N - bits_to_puncture) v = x[i] ? -(float)CONFIRMED_BIT_LLR : (float)CONFIRMED_BIT_LLR;   // :518-526
                if (i >= N - bits_to_puncture) v = 0;                              // :530-533
                if (std::is_same<Q, float>::value) L[i] = (Q)v;
                else {
                    const float s = std::fabs(v) > 20.0f ? q_confirmed : std::min(q_confirmed, std::round(std::fabs(v) * q_scale));
                    L[i] = (Q)(v < 0 ? -s : s);
                }
            }
        }
        decoder.decode_siho(LLRs, dec_bits);                                       // :537
        long bit_errs = 0, frame_errs = 0, iter_sum = 0;
        for (int f = 0; f < frames; ++f) {                                         // m.monitor->check_errors (:538)
            int e = 0;
            for (int i = 0; i < K; ++i) e += dec_bits[(size_t)f * K + i] != ref_bits[(size_t)f * K + i];
            bit_errs += e;
            frame_errs += e > 0;
            iter_sum += decoder.last_iterations()[(size_t)f];
        }
        decoder.reset();                                                           // :540
        std::printf("             %.4f ||  %7d |  %7ld |  %7ld | %.2e | %.2e || mean sweeps %.2f\n", ber, frames, bit_errs, frame_errs,
                    (double)bit_errs / ((double)K * frames), (double)frame_errs / frames, (double)iter_sum / frames);
    }
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 4) {
        std::fprintf(stderr, "Usage: %s filename expansion_factor desired_reconciliation_efficiency [frames_per_ber] [rule[,layered]] "
                             "[f32|i8|i16] [n_ite] [seed]\n", argv[0]);
        return 2;
    }
    int expansion_factor = 0, frames = 100, n_ite = 10;
    float eff = 0;
    if (std::sscanf(argv[2], "%d", &expansion_factor) != 1 || std::sscanf(argv[3], "%f", &eff) != 1 || expansion_factor < 1) {
        std::fprintf(stderr, "Error encountered reading in arguments (are expansion_factor & desired_reconciliation_efficiency numbers?)\n");
        return 2;
    }
    if (argc > 4) frames = std::atoi(argv[4]);
    std::string rule_s = argc > 5 ? argv[5] : "spa", q_s = argc > 6 ? argv[6] : "f32";
    if (argc > 7) n_ite = std::atoi(argv[7]);
    const unsigned seed = argc > 8 ? (unsigned)std::atoi(argv[8]) : 0u;
    module::Schedule sched = module::Schedule::flooding;
    const size_t comma = rule_s.find(',');
    if (comma != std::string::npos) {
        if (rule_s.substr(comma + 1) == "layered") sched = module::Schedule::horizontal_layered;
        rule_s = rule_s.substr(0, comma);
    }
    tools::Update_rule rule = tools::Update_rule_SPA();
    if (rule_s == "ms") rule = tools::Update_rule_MS();
    else if (rule_s.compare(0, 4, "nms:") == 0) rule = tools::Update_rule_NMS((float)std::atof(rule_s.c_str() + 4));
    else if (rule_s.compare(0, 4, "oms:") == 0) rule = tools::Update_rule_OMS((float)std::atof(rule_s.c_str() + 4));
    else if (rule_s != "spa") { std::fprintf(stderr, "unknown rule %s\n", rule_s.c_str()); return 2; }
    try {
        if (q_s == "f32") return run<float>(argv[1], expansion_factor, eff, frames, rule, sched, n_ite, seed, 1.0f, 0.0f);
        if (q_s == "i8") return run<int8_t>(argv[1], expansion_factor, eff, frames, rule, sched, n_ite, seed, 4.0f, 127.0f);
        if (q_s == "i16") return run<int16_t>(argv[1], expansion_factor, eff, frames, rule, sched, n_ite, seed, 64.0f, 2047.0f);
        std::fprintf(stderr, "unknown Q %s\n", q_s.c_str());
        return 2;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 3;
    }
}
