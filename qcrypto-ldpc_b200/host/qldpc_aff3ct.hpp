// AFF3CT-shaped C++ face of the engine, header only, over the C ABI (include/qldpc.h).
//
// The reference drivers hold the decoder as
//     std::unique_ptr<module::Decoder_SISO_SIHO<>> decoder;                       BOOT/src/main.cpp:113
//     decoder = new module::Decoder_LDPC_BP_flooding<B,Q,tools::Update_rule_SPA<Q>>(
//         K, N, n_ite, H, info_bits_pos, tools::Update_rule_SPA<Q>(max_CN_degree),
//         enable_syndrome, syndrome_depth, n_frames);                             BOOT/src/main.cpp:193
// and call  decoder->decode_siho(LLRs, dec_bits)  (:365)  and  decoder->reset()  (:389).
// This class keeps those names, argument meanings (B = one int per bit, Q = the LLR type, n_frames
// frames back to back in one vector) and the error behaviour (exceptions, like tools::invalid_argument /
// tools::length_error / tools::runtime_error), so the driver loops port by changing one line.
// Decoder_SISO_SIHO<B,Q> is templated on Q (BOOT/src/main.cpp:113): Q = float selects the fp32 kernels, Q = int8_t /
// int16_t the fixed-point ones (AFF3CT's Q_8 / Q_16 builds), with Update_rule_NMS's factor restricted to k/8 as in
// AFF3CT's integer normalize<>.
//
// BOOT = errorcorrection/ldpc_examples/my_project_with_aff3ct/examples/bootstrap
#pragma once

#include <cstdint>
#include <memory>
#include <type_traits>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qldpc.h"

namespace qldpc {
namespace tools {

// the update rules the drivers name ("main.cpp (5g-qc)":236-251); LSPA / AMS / Gallager are out of scope
struct Update_rule {
    int rule = QLDPC_RULE_SPA;
    float normalize_factor = 1.0f;   // BOOT/src/main.cpp:98
    float offset = 0.0f;             // BOOT/src/main.cpp:99
};
inline Update_rule Update_rule_SPA(int /*max_check_node_degree*/ = 0) { return Update_rule{QLDPC_RULE_SPA, 1.0f, 0.0f}; }
inline Update_rule Update_rule_MS() { return Update_rule{QLDPC_RULE_NMS, 1.0f, 0.0f}; }
inline Update_rule Update_rule_NMS(float normalize_factor) { return Update_rule{QLDPC_RULE_NMS, normalize_factor, 0.0f}; }
inline Update_rule Update_rule_OMS(float offset) { return Update_rule{QLDPC_RULE_OMS, 1.0f, offset}; }

struct invalid_argument : std::invalid_argument { using std::invalid_argument::invalid_argument; };
struct length_error : std::length_error { using std::length_error::length_error; };
struct runtime_error : std::runtime_error { using std::runtime_error::runtime_error; };

// parity-check matrix handle: tools::LDPC_matrix_handler::read(...) ("main.cpp (alist)":340, "(5g-qc)":389)
class Sparse_matrix {
public:
    static Sparse_matrix read(const std::string &path)
    {
        qldpc_code *c = nullptr;
        const bool qc = path.size() > 3 && path.compare(path.size() - 3, 3, ".qc") == 0;
        const int rc = qc ? qldpc_code_from_qc_file(path.c_str(), &c) : qldpc_code_from_alist_file(path.c_str(), &c);
        if (rc != QLDPC_OK) throw runtime_error("LDPC_matrix_handler::read(" + path + "): " + qldpc_strerror(rc));
        return Sparse_matrix(c);
    }
    const qldpc_code *get() const { return code_.get(); }
    qldpc_code *get() { return code_.get(); }
    qldpc_code_info info() const
    {
        qldpc_code_info i{};
        qldpc_code_get_info(code_.get(), &i);
        return i;
    }
    // H is stored transposed in AFF3CT (rows = variables): get_n_rows() = N, get_n_cols() = M,
    // get_cols_max_degree() = max check degree (BOOT/src/main.cpp:178)
    int get_n_rows() const { return info().n; }
    int get_n_cols() const { return info().m; }
    int get_cols_max_degree() const { return info().max_chk_degree; }

private:
    explicit Sparse_matrix(qldpc_code *c) : code_(c, &qldpc_code_free) {}
    std::shared_ptr<qldpc_code> code_;
};

}  // namespace tools

namespace module {

enum class Schedule { flooding = QLDPC_SCHED_FLOODING, horizontal_layered = QLDPC_SCHED_LAYERED };

// Decoder_SISO_SIHO<B,Q>-shaped decoder running on the GPU.
template <typename B = int, typename Q = float>
class Decoder_LDPC_BP {
public:
    Decoder_LDPC_BP(int K, int N, int n_ite, tools::Sparse_matrix H, const std::vector<uint32_t> &info_bits_pos,
                    const tools::Update_rule &up_rule, bool enable_syndrome = true, int syndrome_depth = 1,
                    int n_frames = 1, Schedule schedule = Schedule::flooding, int device = 0)
        : K_(K), N_(N), n_frames_(n_frames), H_(std::move(H))
    {
        const qldpc_code_info inf = H_.info();
        if (N != inf.n) throw tools::invalid_argument("'N' has to be equal to 'H.get_n_rows()'");
        if (K <= 0 || K > N) throw tools::invalid_argument("'K' has to be in ]0, N]");
        if (n_ite <= 0) throw tools::invalid_argument("'n_ite' has to be greater than 0");
        if (n_frames <= 0) throw tools::invalid_argument("'n_frames' has to be greater than 0");
        if (syndrome_depth <= 0) throw tools::invalid_argument("'syndrome_depth' has to be greater than 0");
        if ((int)info_bits_pos.size() != K) throw tools::length_error("'info_bits_pos.size()' has to be equal to 'K'");
        std::vector<int32_t> pos(info_bits_pos.begin(), info_bits_pos.end());
        int rc = qldpc_code_set_info_bits_pos(H_.get(), pos.data(), K);
        if (rc != QLDPC_OK) throw tools::invalid_argument("'info_bits_pos' holds a position outside [0, N[");
        qldpc_decoder_config cfg;
        qldpc_decoder_config_default(&cfg);
        cfg.schedule = (int)schedule;
        cfg.rule = up_rule.rule;
        // every decoder the reference instantiates is B=int, Q=float (BOOT/src/main.cpp:113); the integer Q types reach the
        // fixed-point kernels (BG1 Z=384 layered int8: the streamed kernel of the headline benchmark)
        static_assert(std::is_same<Q, float>::value || std::is_same<Q, int8_t>::value || std::is_same<Q, int16_t>::value,
                      "Q has to be float, int8_t or int16_t");
        cfg.dtype = std::is_same<Q, float>::value ? QLDPC_DTYPE_F32 : (std::is_same<Q, int8_t>::value ? QLDPC_DTYPE_I8 : QLDPC_DTYPE_I16);
        if (cfg.dtype != QLDPC_DTYPE_F32 && up_rule.rule == QLDPC_RULE_SPA)
            throw tools::invalid_argument("Update_rule_SPA needs a floating-point 'Q'");
        cfg.max_iter = n_ite;
        cfg.early_stop = enable_syndrome ? 1 : 0;
        cfg.syndrome_depth = syndrome_depth;
        cfg.norm_factor = up_rule.normalize_factor;
        cfg.offset = up_rule.offset;
        cfg.out_mode = QLDPC_OUT_INFO;
        cfg.device = device;
        qldpc_decoder *d = nullptr;
        rc = qldpc_decoder_create(H_.get(), &cfg, &d);
        if (rc == QLDPC_ERR_ARG && cfg.dtype != QLDPC_DTYPE_F32)
            throw tools::invalid_argument("integer 'Q': 'normalize_factor' has to be k/8 (k = 1..8), 'offset' a non-negative integer");
        if (rc != QLDPC_OK) throw tools::runtime_error(std::string("qldpc_decoder_create: ") + qldpc_strerror(rc));
        dec_.reset(d, &qldpc_decoder_free);
        out_words_ = qldpc_out_words(d);
    }

    int get_K() const { return K_; }
    int get_N() const { return N_; }
    int get_n_frames() const { return n_frames_; }

    // decode_siho(Y_N, V_K): n_frames*N LLRs in, n_frames*K hard decisions out, one B per bit (BOOT/src/main.cpp:365)
    void decode_siho(const std::vector<Q> &Y_N, std::vector<B> &V_K)
    {
        check_sizes(Y_N.size(), (size_t)N_, "Y_N");
        check_sizes(V_K.size(), (size_t)K_, "V_K");
        packed_.assign((size_t)n_frames_ * out_words_, 0u);
        ok_.assign((size_t)n_frames_, 0);
        iters_.assign((size_t)n_frames_, 0);
        const int rc = qldpc_decode(dec_.get(), Y_N.data(), nullptr, n_frames_, packed_.data(), ok_.data(), iters_.data(), nullptr);
        if (rc != QLDPC_OK) throw tools::runtime_error(std::string("decode_siho: ") + qldpc_strerror(rc));
        for (int f = 0; f < n_frames_; ++f)
            for (int i = 0; i < K_; ++i)
                V_K[(size_t)f * K_ + i] = (B)((packed_[(size_t)f * out_words_ + i / 32] >> (31 - (i % 32))) & 1u);
    }

    // decode_siso(Y_N1, Y_N2): a-posteriori LLRs of all N positions
    void decode_siso(const std::vector<Q> &Y_N1, std::vector<Q> &Y_N2)
    {
        check_sizes(Y_N1.size(), (size_t)N_, "Y_N1");
        check_sizes(Y_N2.size(), (size_t)N_, "Y_N2");
        // a-posteriori values come back as float (Q = float) or int32 (integer Q, saturated to Q's range here)
        typedef typename std::conditional<std::is_same<Q, float>::value, float, int32_t>::type P;
        std::vector<P> post((size_t)n_frames_ * N_);
        packed_.assign((size_t)n_frames_ * out_words_, 0u);
        ok_.assign((size_t)n_frames_, 0);
        iters_.assign((size_t)n_frames_, 0);
        const int rc = qldpc_decode(dec_.get(), Y_N1.data(), nullptr, n_frames_, packed_.data(), ok_.data(), iters_.data(), post.data());
        if (rc != QLDPC_OK) throw tools::runtime_error(std::string("decode_siso: ") + qldpc_strerror(rc));
        for (size_t i = 0; i < post.size(); ++i) {
            if (std::is_same<Q, float>::value) Y_N2[i] = (Q)post[i];
            else {
                const long lo = std::is_same<Q, int8_t>::value ? -128 : -32768, hi = -lo - 1;
                const long v = (long)post[i];
                Y_N2[i] = (Q)(v < lo ? lo : (v > hi ? hi : v));
            }
        }
    }

    // AFF3CT keeps check-to-variable messages between calls until reset(); this engine starts every call
    // from zero messages, which is what the drivers get by calling reset() after each frame (BOOT/src/main.cpp:389).
    void reset() {}

    // extras the AFF3CT class does not have
    const std::vector<uint8_t> &last_syndrome_ok() const { return ok_; }
    const std::vector<uint16_t> &last_iterations() const { return iters_; }
    const char *kernel_name() const { return qldpc_decoder_kernel_name(dec_.get()); }

private:
    void check_sizes(size_t got, size_t per_frame, const char *name) const
    {
        if (got != per_frame * (size_t)n_frames_)
            throw tools::length_error(std::string("'") + name + ".size()' has to be equal to '" + std::to_string(per_frame) +
                                      "' * 'n_frames' ('" + name + ".size()' = " + std::to_string(got) + ")");
    }
    int K_, N_, n_frames_, out_words_ = 0;
    tools::Sparse_matrix H_;
    std::shared_ptr<qldpc_decoder> dec_;
    std::vector<uint32_t> packed_;
    std::vector<uint8_t> ok_;
    std::vector<uint16_t> iters_;
};

// Encoder_LDPC_from_QC<B> ("main.cpp (5g-qc)":177): systematic encoder of a 5G-NR shaped quasi-cyclic code, n_frames
// frames per call, one B per bit; info_bits_pos = [0, K) (get_info_bits_pos()).
template <typename B = int>
class Encoder_LDPC_from_QC {
public:
    Encoder_LDPC_from_QC(int K, int N, tools::Sparse_matrix H, int n_frames = 1, int device = 0)
        : K_(K), N_(N), n_frames_(n_frames), H_(std::move(H))
    {
        const qldpc_code_info inf = H_.info();
        if (N != inf.n || K != inf.n - inf.m) throw tools::invalid_argument("'K' / 'N' do not match H");
        qldpc_decoder_config cfg;
        qldpc_decoder_config_default(&cfg);
        cfg.max_iter = 1;
        cfg.device = device;
        qldpc_decoder *d = nullptr;
        const int rc = qldpc_decoder_create(H_.get(), &cfg, &d);
        if (rc != QLDPC_OK) throw tools::runtime_error(std::string("qldpc_decoder_create: ") + qldpc_strerror(rc));
        dec_.reset(d, &qldpc_decoder_free);
    }
    std::vector<uint32_t> get_info_bits_pos() const
    {
        std::vector<uint32_t> p((size_t)K_);
        for (int i = 0; i < K_; ++i) p[(size_t)i] = (uint32_t)i;
        return p;
    }
    void encode(const std::vector<B> &U_K, std::vector<B> &X_N)
    {
        if (U_K.size() != (size_t)K_ * n_frames_ || X_N.size() != (size_t)N_ * n_frames_)
            throw tools::length_error("'U_K.size()' / 'X_N.size()' have to be 'K' / 'N' * 'n_frames'");
        const int kw = (K_ + 31) / 32, cw = qldpc_codeword_words(dec_.get());
        std::vector<uint32_t> msg((size_t)n_frames_ * kw, 0u), cword((size_t)n_frames_ * cw);
        for (int f = 0; f < n_frames_; ++f)
            for (int i = 0; i < K_; ++i)
                if (U_K[(size_t)f * K_ + i]) msg[(size_t)f * kw + i / 32] |= 1u << (31 - i % 32);
        const int rc = qldpc_encode_nr(dec_.get(), msg.data(), n_frames_, cword.data());
        if (rc != QLDPC_OK) throw tools::runtime_error(std::string("Encoder_LDPC_from_QC::encode: ") + qldpc_strerror(rc));
        for (int f = 0; f < n_frames_; ++f)
            for (int i = 0; i < N_; ++i) X_N[(size_t)f * N_ + i] = (B)((cword[(size_t)f * cw + i / 32] >> (31 - i % 32)) & 1u);
    }

private:
    int K_, N_, n_frames_;
    tools::Sparse_matrix H_;
    std::shared_ptr<qldpc_decoder> dec_;
};

// Encoder_LDPC<B>(K, N, G, n_frames) ("main.cpp (alist)":143, encode at :417) and Encoder_LDPC_from_H<B>(K, N, H, ...)
// ("main.cpp (alist-v1.0.1)":144): systematic encoders of an arbitrary code -- from a generator matrix file (the reference
// reads G with LDPC_matrix_handler::read and hands the Sparse_matrix over; here the file path is the argument), or from H by
// elimination (AFF3CT's "IDENTITY" method).  get_info_bits_pos() is what the reference passes on to the decoder.
template <typename B = int>
class Encoder_LDPC {
public:
    Encoder_LDPC(int K, int N, const std::string &G_alist_path, int n_frames = 1, int device = 0) : n_frames_(n_frames)
    {
        qldpc_encoder *e = nullptr;
        const int rc = qldpc_encoder_from_g_alist_file(G_alist_path.c_str(), device, &e);
        if (rc != QLDPC_OK) throw tools::runtime_error(std::string("Encoder_LDPC: ") + qldpc_strerror(rc));
        adopt(e, K, N);
    }
    std::vector<uint32_t> get_info_bits_pos() const
    {
        std::vector<int32_t> p((size_t)K_);
        qldpc_encoder_info_bits_pos(enc_.get(), p.data());
        return std::vector<uint32_t>(p.begin(), p.end());
    }
    void encode(const std::vector<B> &U_K, std::vector<B> &X_N)
    {
        if (U_K.size() != (size_t)K_ * n_frames_ || X_N.size() != (size_t)N_ * n_frames_)
            throw tools::length_error("'U_K.size()' / 'X_N.size()' have to be 'K' / 'N' * 'n_frames'");
        const int kw = (K_ + 31) / 32, cw = (N_ + 31) / 32;
        std::vector<uint32_t> msg((size_t)n_frames_ * kw, 0u), cword((size_t)n_frames_ * cw);
        for (int f = 0; f < n_frames_; ++f)
            for (int i = 0; i < K_; ++i)
                if (U_K[(size_t)f * K_ + i]) msg[(size_t)f * kw + i / 32] |= 1u << (31 - i % 32);
        const int rc = qldpc_encode(enc_.get(), msg.data(), n_frames_, cword.data());
        if (rc != QLDPC_OK) throw tools::runtime_error(std::string("Encoder_LDPC::encode: ") + qldpc_strerror(rc));
        for (int f = 0; f < n_frames_; ++f)
            for (int i = 0; i < N_; ++i) X_N[(size_t)f * N_ + i] = (B)((cword[(size_t)f * cw + i / 32] >> (31 - i % 32)) & 1u);
    }

protected:
    explicit Encoder_LDPC(int n_frames) : n_frames_(n_frames) {}
    void adopt(qldpc_encoder *e, int K, int N)
    {
        enc_.reset(e, &qldpc_encoder_free);
        int32_t k = 0, n = 0;
        qldpc_encoder_get_info(e, &k, &n);
        if (K != k || N != n) throw tools::invalid_argument("'K' / 'N' do not match the matrix");
        K_ = K; N_ = N;
    }
    int K_ = 0, N_ = 0, n_frames_;
    std::shared_ptr<qldpc_encoder> enc_;
};

template <typename B = int>
class Encoder_LDPC_from_H : public Encoder_LDPC<B> {
public:
    Encoder_LDPC_from_H(int K, int N, const tools::Sparse_matrix &H, int n_frames = 1, int device = 0) : Encoder_LDPC<B>(n_frames)
    {
        qldpc_encoder *e = nullptr;
        const int rc = qldpc_encoder_from_h(H.get(), device, &e);
        if (rc != QLDPC_OK) throw tools::runtime_error(std::string("Encoder_LDPC_from_H: ") + qldpc_strerror(rc));
        this->adopt(e, K, N);
    }
};

}  // namespace module
}  // namespace qldpc
