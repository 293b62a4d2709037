// Blind (incremental-redundancy) LDPC reconciliation for ecd2, on top of the C ABI of libqldpc_b200.
//
// What it replaces: the Cascade/BICONF packet ping-pong of ecd2 (EC/subcomponents/cascade_biconf.c:427-940) for the
// two algorithm slots the reference reserves but never implemented: ALG_LDPC_CONTINUE_ROLES / ALG_LDPC_FLIP_ROLES
// (EC/definitions/algorithms/algorithms.h:36-42; both dispatch sites return error 81, EC/subcomponents/qber_estim.c:
// 337-340,420-423).  The classes below are the bodies of the packet handlers a maintainer registers
// (PacketHandlerArray, EC/definitions/algorithms/packet_manager.h:33): each consumes one received packet
// (`char *receivebuf`) and returns the packets to queue with comms_insertSendPacket (EC/subcomponents/comms.c:16-38).
//
// Protocol (the reference's send-parity formulation, EC/README_AFF3CT.md:42, BOOT/src/main.cpp:351-354, with 5G-NR
// rate matching as in ML/BPSK_nrldpc_sim_RM_FP.m:16-21 -- a prefix of the parity block columns is transmitted):
//   Alice (EC initiator)                              Bob (EC follower)
//   cuts the block into frames of K = 22 Z key bits,
//   NR-encodes them, sends the first m0(QBER) parity
//   block rows                 -- subtype 9  LDPC_PARITY -->   decodes with the rate-matched code (rows 0..m0-1)
//                              <-- subtype 10 LDPC_NACK  --    lists the frames whose syndrome check failed
//   sends the next DELTA parity rows of those frames
//   (or, after the last row, the key bits themselves)
//                              -- subtype 11 LDPC_MORE  -->    decodes those frames again with more rows
//                              <-- subtype 12 LDPC_DONE  --    corrected bits are in the block; carries one CRC-32 per frame
//   compares them with the CRCs of her own frames (the confirmation step the reference planned, EC/README_LDPC.md:784-788);
//   a frame whose CRC differs is revealed (subtype 11 with reveal=1) and Bob answers with a new LDPC_DONE;
//   when all agree both sides call privAmp_sendPrivAmpMsgAndPrivAmp (cascade_biconf.c:892)
// Leakage: every parity bit sent counts once in pb->leakageBits (EC/subcomponents/priv_amp.c:47,166), a revealed frame
// counts its K key bits, and the confirmation counts 32 bits per frame the first time the frame's CRC-32 crosses the
// channel (CRC-32 is affine in the key: 32 linear parities, exactly like the parity bits cascade_biconf.c counts with
// leakageBits++).  Both sides keep the same count (Bob's KeyBlock.leakageBits mirrors Alice's).
// Received packets are not trusted: every length, index and rate field is checked against totalLengthInBytes and the
// block before anything is copied; a violation returns an ecd2 error code (the reference handlers do the same, e.g.
// qber_estim.c:372 returns 52).  Bit vectors are MSB-first 32-bit words (EC/subcomponents/helpers.h:65-68), packets start with
// EcPktHdr_Base (EC/definitions/packets.h:65-71), host-endian, below transferd's 10 000-byte cap per frame
// (remotecrypto/transferd.h:139): a packet carries at most `frames_per_packet` frames.
//
// One handler call may serve a BATCH of process blocks (all blocks whose packets are pending when the select loop
// wakes up): their frames go to the GPU in one launch.  There is no CPU decoder behind this: without an sm_100 device
// decoder creation fails and the handler returns the ecd2 error code.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qldpc.h"

namespace qldpc {
namespace ecd2 {

constexpr uint32_t EC_PACKET_TAG = 6;   // EC/definitions/packets.h:57
enum : uint32_t {                       // continues EC_SUBTYPES (packets.h:46-55), which ends at 8
    SUBTYPE_LDPC_PARITY = 9,
    SUBTYPE_LDPC_NACK = 10,
    SUBTYPE_LDPC_MORE = 11,
    SUBTYPE_LDPC_DONE = 12,
};
constexpr int ERR_LDPC_UNSUPPORTED = 81;   // errormessage[81] "Unsupported functionality" (EC/ecd2.h:333)
constexpr int ERR_LDPC_NO_BLOCK = 49;      // errormessage[49] "cannot find processBlock in list"
constexpr int ERR_LDPC_BAD_PACKET = 85;    // appended by integration/ecd2_ldpc.patch: "LDPC packet inconsistent with its block"

struct EcPktHdr_Base {                  // EC/definitions/packets.h:65-71
    uint32_t tag, totalLengthInBytes, subtype, epoch, numberOfEpochs;
};
struct EcPktHdr_LdpcParity {            // followed by frames * rows * z/32 parity words
    EcPktHdr_Base base;
    uint32_t z, frames, first_frame, rows;   // rows: parity block rows sent (0 .. rows-1)
    uint32_t workbits;
    float qber;
};
struct EcPktHdr_LdpcNack {              // followed by n_failed frame indices (block-relative)
    EcPktHdr_Base base;
    uint32_t round, n_failed;
};
struct EcPktHdr_LdpcMore {              // followed by n_frames indices, then per frame (row_to-row_from)*z/32 parity words,
    EcPktHdr_Base base;                 // or, if reveal != 0, the frame's 22*z/32 key words
    uint32_t round, n_frames, row_from, row_to, reveal;
};
struct EcPktHdr_LdpcDone {              // followed by `frames` CRC-32 values (qldpc_crc32_frames over the K key bits of a frame)
    EcPktHdr_Base base;
    uint32_t rounds, frames_revealed, frames;
    uint32_t corrected_errors;          // Bob's pb->correctedErrors: Alice needs it for the error-rate estimate of priv_amp.c:118
};

// the ProcessBlock fields an LDPC handler touches (EC/definitions/processblock.h:102-129)
struct KeyBlock {
    uint32_t startEpoch = 0, numberOfEpochs = 1;
    uint32_t *mainBufPtr = nullptr;     // key bits, MSB-first words; corrected in place on Bob's side
    int workbits = 0;
    float localError = 0.f;             // estimated QBER
    int leakageBits = 0, correctedErrors = 0;
};

typedef std::vector<uint8_t> Packet;

struct Params {
    std::string base_qc;                // NR base graph file (.qc: "cols rows Z" + shift table), BG1: 68 x 46
    float f_start = 1.25f;              // initial efficiency target: m0 = ceil(f_start * 22 * h(QBER)), at least 4 rows
    int delta_rows = 2;                 // extra parity block rows per NACK round
    int max_iter = 20;
    float norm_factor = 0.75f;
    int device = 0;
    int frames_per_packet = 4;          // 4 * 46 * 384 / 8 = 8832 B < transferd's 10 000-byte EC packet cap
};

inline double h2(double p) { return (p <= 0 || p >= 1) ? 0.0 : -p * std::log2(p) - (1 - p) * std::log2(1 - p); }

// ---- the rate-compatible code family: row prefixes of the NR base graph -------------------------------------------
class CodeFamily {
public:
    explicit CodeFamily(const Params &p) : prm(p)
    {
        std::ifstream in(p.base_qc);
        if (!in) throw std::runtime_error("qldpc_blind: cannot read " + p.base_qc);
        in >> cols >> rows >> z;
        base.resize((size_t)rows * cols);
        for (auto &v : base) in >> v;
        if (!in || rows < 4 || cols - rows < 1 || z % 32) throw std::runtime_error("qldpc_blind: bad base graph file");
        kcols = cols - rows;
    }
    ~CodeFamily()
    {
        for (auto &kv : decs) qldpc_decoder_free(kv.second);
        for (auto &kv : codes) qldpc_code_free(kv.second);
    }
    int K() const { return kcols * z; }
    int kwords() const { return K() / 32; }
    int zwords() const { return z / 32; }
    int max_rows() const { return rows; }
    int initial_rows(float qber) const
    {
        const int m = (int)std::ceil(prm.f_start * kcols * h2(qber));
        return std::min(rows, std::max(4, m));
    }
    // decoder of the code made of the first m block rows and the first kcols + m block columns
    qldpc_decoder *decoder(int m)
    {
        auto it = decs.find(m);
        if (it != decs.end()) return it->second;
        const int c = kcols + m;
        std::vector<int32_t> sub((size_t)m * c);
        for (int r = 0; r < m; ++r)
            for (int j = 0; j < c; ++j) sub[(size_t)r * c + j] = base[(size_t)r * cols + j] < 0 ? -1 : base[(size_t)r * cols + j] % z;
        qldpc_code *code = nullptr;
        if (qldpc_code_from_qc(sub.data(), m, c, z, &code)) throw std::runtime_error("qldpc_blind: qldpc_code_from_qc failed");
        codes[m] = code;
        qldpc_decoder_config cfg;
        qldpc_decoder_config_default(&cfg);
        cfg.schedule = QLDPC_SCHED_LAYERED; cfg.rule = QLDPC_RULE_NMS; cfg.dtype = QLDPC_DTYPE_I8;
        cfg.norm_factor = prm.norm_factor; cfg.max_iter = prm.max_iter; cfg.early_stop = 1;
        cfg.out_mode = QLDPC_OUT_INFO; cfg.device = prm.device;
        qldpc_decoder *d = nullptr;
        const int rc = qldpc_decoder_create(code, &cfg, &d);
        if (rc) throw std::runtime_error(std::string("qldpc_blind: qldpc_decoder_create: ") + qldpc_strerror(rc));
        decs[m] = d;
        return d;
    }
    Params prm;
    int cols = 0, rows = 0, z = 0, kcols = 0;
    std::vector<int32_t> base;

private:
    std::map<int, qldpc_code *> codes;
    std::map<int, qldpc_decoder *> decs;
};

namespace detail {
template <class H>
Packet make_packet(uint32_t subtype, const KeyBlock &b, const H &hdr_fields, size_t payload_bytes)
{
    Packet p(sizeof(H) + payload_bytes);
    H h = hdr_fields;
    h.base.tag = EC_PACKET_TAG;
    h.base.totalLengthInBytes = (uint32_t)p.size();
    h.base.subtype = subtype;
    h.base.epoch = b.startEpoch;
    h.base.numberOfEpochs = b.numberOfEpochs;
    std::memcpy(p.data(), &h, sizeof(H));
    return p;
}
inline int frames_of(const KeyBlock &b, int K) { return (b.workbits + K - 1) / K; }
// header H at the start of a received packet whose declared length must be exactly sizeof(H) + payload(H) bytes
template <class H>
bool read_header(const char *buf, uint32_t subtype, H &h)
{
    std::memcpy(&h, buf, sizeof(H));
    return h.base.tag == EC_PACKET_TAG && h.base.subtype == subtype && h.base.totalLengthInBytes >= sizeof(H);
}
inline bool payload_is(const EcPktHdr_Base &b, size_t header, uint64_t payload) { return (uint64_t)b.totalLengthInBytes == header + payload; }
// key bits of frame f of a block as kwords words (the last frame is zero padded on both sides)
inline void copy_frame_key(const KeyBlock &b, int f, int kwords, uint32_t *dst)
{
    const int total_words = (b.workbits + 31) / 32;
    for (int w = 0; w < kwords; ++w) {
        const int src = f * kwords + w;
        uint32_t v = src < total_words ? b.mainBufPtr[src] : 0u;
        if (src == total_words - 1 && (b.workbits & 31)) v &= ~0u << (32 - (b.workbits & 31));   // bits beyond workbits
        dst[w] = v;
    }
}
// CRC-32 of every frame of a block (K key bits each, the last frame zero padded), on the GPU
inline int block_crcs(const KeyBlock &b, int K, int kwords, int device, std::vector<uint32_t> &crc)
{
    const int nf = frames_of(b, K);
    std::vector<uint32_t> buf((size_t)nf * kwords);
    for (int f = 0; f < nf; ++f) copy_frame_key(b, f, kwords, buf.data() + (size_t)f * kwords);
    crc.assign(nf, 0u);
    return qldpc_crc32_frames(device, buf.data(), nf, kwords, kwords, crc.data());
}
}  // namespace detail

// ---- Alice: EC initiator --------------------------------------------------------------------------------------------
class BlindAlice {
public:
    explicit BlindAlice(std::shared_ptr<CodeFamily> fam) : fam_(std::move(fam)) {}

    // body of the `case ALG_LDPC_*` in qber_prepareErrorCorrection (qber_estim.c:420-423): encode, send the first rows
    int initiate(std::vector<KeyBlock *> &blocks, std::vector<Packet> &send)
    {
        const int K = fam_->K(), kw = fam_->kwords(), zw = fam_->zwords();
        int F = 0;
        for (auto *b : blocks) F += detail::frames_of(*b, K);
        std::vector<uint32_t> msg((size_t)F * kw);
        int f0 = 0;
        for (auto *b : blocks) {
            const int nf = detail::frames_of(*b, K);
            for (int f = 0; f < nf; ++f) detail::copy_frame_key(*b, f, kw, msg.data() + (size_t)(f0 + f) * kw);
            f0 += nf;
        }
        qldpc_decoder *full = fam_->decoder(fam_->max_rows());
        const int cw = qldpc_codeword_words(full);
        std::vector<uint32_t> cword((size_t)F * cw);
        if (qldpc_encode_nr(full, msg.data(), F, cword.data())) return ERR_LDPC_UNSUPPORTED;
        f0 = 0;
        for (auto *b : blocks) {
            State &s = st_[b->startEpoch];
            s.frames = detail::frames_of(*b, K);
            s.parity.assign((size_t)s.frames * fam_->max_rows() * zw, 0u);
            for (int f = 0; f < s.frames; ++f)
                std::memcpy(s.parity.data() + (size_t)f * fam_->max_rows() * zw, cword.data() + (size_t)(f0 + f) * cw + kw,
                            (size_t)fam_->max_rows() * zw * 4);
            f0 += s.frames;
            s.rows_sent.assign(s.frames, fam_->initial_rows(b->localError));
            s.revealed.assign(s.frames, 0);
            s.crc_seen.assign(s.frames, 0);
            const int m0 = s.rows_sent[0];
            for (int first = 0; first < s.frames; first += fam_->prm.frames_per_packet) {
                const int n = std::min(fam_->prm.frames_per_packet, s.frames - first);
                EcPktHdr_LdpcParity h{};
                h.z = (uint32_t)fam_->z; h.frames = (uint32_t)n; h.first_frame = (uint32_t)first; h.rows = (uint32_t)m0;
                h.workbits = (uint32_t)b->workbits; h.qber = b->localError;
                Packet p = detail::make_packet(SUBTYPE_LDPC_PARITY, *b, h, (size_t)n * m0 * zw * 4);
                for (int f = 0; f < n; ++f)
                    std::memcpy(p.data() + sizeof(h) + (size_t)f * m0 * zw * 4,
                                s.parity.data() + (size_t)(first + f) * fam_->max_rows() * zw, (size_t)m0 * zw * 4);
                send.push_back(std::move(p));
            }
            b->leakageBits += s.frames * m0 * fam_->z;
        }
        return 0;
    }

    // handler for SUBTYPE_LDPC_NACK: send DELTA more parity rows of the listed frames (or the key bits after the last row).
    // Frames of one NACK may sit at different rates (failures of several decode groups are answered together): they are
    // grouped by the rows already sent, one MORE packet per (from, to) and at most frames_per_packet frames.
    int on_nack(KeyBlock &b, const char *receivebuf, std::vector<Packet> &send)
    {
        EcPktHdr_LdpcNack in;
        if (!detail::read_header(receivebuf, SUBTYPE_LDPC_NACK, in)) return ERR_LDPC_BAD_PACKET;
        auto sit = st_.find(b.startEpoch);
        if (sit == st_.end()) return ERR_LDPC_NO_BLOCK;
        State &s = sit->second;
        if (in.n_failed == 0 || in.n_failed > (uint32_t)s.frames || !detail::payload_is(in.base, sizeof(in), 4ull * in.n_failed))
            return ERR_LDPC_BAD_PACKET;
        const uint32_t *idx = reinterpret_cast<const uint32_t *>(receivebuf + sizeof(in));
        const int zw = fam_->zwords(), kw = fam_->kwords(), R = fam_->max_rows();
        std::map<int, std::vector<uint32_t>> by_rate;
        for (uint32_t k = 0; k < in.n_failed; ++k) {
            if (idx[k] >= (uint32_t)s.frames) return ERR_LDPC_BAD_PACKET;
            by_rate[s.rows_sent[idx[k]]].push_back(idx[k]);
        }
        for (auto &kv : by_rate) {
            const int from = kv.first;
            const std::vector<uint32_t> &fr = kv.second;
            const bool reveal = from >= R;
            const int to = reveal ? R : std::min(R, from + fam_->prm.delta_rows);
            const size_t per = reveal ? (size_t)kw * 4 : (size_t)(to - from) * zw * 4;
            for (size_t first = 0; first < fr.size(); first += (size_t)fam_->prm.frames_per_packet) {
                const uint32_t n = (uint32_t)std::min<size_t>((size_t)fam_->prm.frames_per_packet, fr.size() - first);
                EcPktHdr_LdpcMore h{};
                h.round = in.round; h.n_frames = n; h.row_from = (uint32_t)from; h.row_to = (uint32_t)to; h.reveal = reveal;
                Packet p = detail::make_packet(SUBTYPE_LDPC_MORE, b, h, (size_t)n * 4 + n * per);
                std::memcpy(p.data() + sizeof(h), fr.data() + first, (size_t)n * 4);
                for (uint32_t k = 0; k < n; ++k) {
                    const int f = (int)fr[first + k];
                    uint8_t *dst = p.data() + sizeof(h) + (size_t)n * 4 + k * per;
                    if (reveal) {
                        std::vector<uint32_t> key(kw);
                        detail::copy_frame_key(b, f, kw, key.data());
                        std::memcpy(dst, key.data(), per);
                        if (!s.revealed[f]) { b.leakageBits += fam_->K(); s.revealed[f] = 1; }
                    } else {
                        std::memcpy(dst, s.parity.data() + ((size_t)f * R + from) * zw, per);
                        b.leakageBits += (to - from) * fam_->z;
                        s.rows_sent[f] = to;
                    }
                }
                send.push_back(std::move(p));
            }
        }
        return 0;
    }

    // handler for SUBTYPE_LDPC_DONE: compare Bob's frame CRCs with her own.  All equal: the block is reconciled (`confirmed`
    // is set; next: privacy amplification).  Otherwise the differing frames are revealed and Bob will send a new DONE.
    int on_done(KeyBlock &b, const char *receivebuf, std::vector<Packet> &send, bool &confirmed)
    {
        EcPktHdr_LdpcDone in;
        if (!detail::read_header(receivebuf, SUBTYPE_LDPC_DONE, in)) return ERR_LDPC_BAD_PACKET;
        auto sit = st_.find(b.startEpoch);
        if (sit == st_.end()) return ERR_LDPC_NO_BLOCK;
        State &s = sit->second;
        if (in.frames != (uint32_t)s.frames || !detail::payload_is(in.base, sizeof(in), 4ull * in.frames)) return ERR_LDPC_BAD_PACKET;
        const uint32_t *theirs = reinterpret_cast<const uint32_t *>(receivebuf + sizeof(in));
        std::vector<uint32_t> mine;
        if (detail::block_crcs(b, fam_->K(), fam_->kwords(), fam_->prm.device, mine)) return ERR_LDPC_UNSUPPORTED;
        // the CRC-32 of a frame is 32 linear parities of its key bits in the clear: counted once per frame (a frame that had
        // been revealed before has nothing left to leak)
        for (int f = 0; f < s.frames; ++f)
            if (!s.crc_seen[f]) { s.crc_seen[f] = 1; if (!s.revealed[f]) b.leakageBits += 32; }
        std::vector<uint32_t> bad;
        for (uint32_t f = 0; f < in.frames && f < mine.size(); ++f)
            if (mine[f] != theirs[f]) bad.push_back(f);
        confirmed = bad.empty() && in.frames == mine.size();
        if (confirmed) { b.correctedErrors = (int)in.corrected_errors; st_.erase(b.startEpoch); return 0; }
        mismatches_ += (int)bad.size();
        const int kw = fam_->kwords();
        for (size_t first = 0; first < bad.size(); first += (size_t)fam_->prm.frames_per_packet) {
            const uint32_t n = (uint32_t)std::min<size_t>((size_t)fam_->prm.frames_per_packet, bad.size() - first);
            EcPktHdr_LdpcMore h{};
            h.round = in.rounds + 1; h.n_frames = n; h.row_from = h.row_to = (uint32_t)fam_->max_rows(); h.reveal = 1;
            Packet p = detail::make_packet(SUBTYPE_LDPC_MORE, b, h, (size_t)n * 4 + (size_t)n * kw * 4);
            std::memcpy(p.data() + sizeof(h), bad.data() + first, (size_t)n * 4);
            for (uint32_t k = 0; k < n; ++k) {
                std::vector<uint32_t> key(kw);
                detail::copy_frame_key(b, (int)bad[first + k], kw, key.data());
                std::memcpy(p.data() + sizeof(h) + (size_t)n * 4 + (size_t)k * kw * 4, key.data(), (size_t)kw * 4);
                if (!s.revealed[bad[first + k]]) { b.leakageBits += fam_->K() - 32; s.revealed[bad[first + k]] = 1; }   // its CRC was counted
            }
            send.push_back(std::move(p));
        }
        return 0;
    }
    int mismatches() const { return mismatches_; }

private:
    struct State {
        int frames = 0;
        std::vector<uint32_t> parity;
        std::vector<int> rows_sent;
        std::vector<char> revealed, crc_seen;   // leakage is counted once per revealed frame / disclosed CRC
    };
    std::shared_ptr<CodeFamily> fam_;
    std::map<uint32_t, State> st_;
    int mismatches_ = 0;
};

// ---- Bob: EC follower -----------------------------------------------------------------------------------------------
class BlindBob {
public:
    explicit BlindBob(std::shared_ptr<CodeFamily> fam) : fam_(std::move(fam)) {}

    // handler for SUBTYPE_LDPC_PARITY.  Parity packets of several blocks may be handed over together (`pkts[i]` belongs to
    // `blocks[i]`; a block appears once per packet): everything pending is decoded in one launch per rate.
    int on_parity(const std::vector<KeyBlock *> &blocks, const std::vector<const char *> &pkts, std::vector<Packet> &send)
    {
        const int zw = fam_->zwords(), R = fam_->max_rows(), K = fam_->K();
        for (size_t i = 0; i < pkts.size(); ++i) {
            EcPktHdr_LdpcParity in;
            if (!detail::read_header(pkts[i], SUBTYPE_LDPC_PARITY, in)) return ERR_LDPC_BAD_PACKET;
            KeyBlock &b = *blocks[i];
            const int nf = detail::frames_of(b, K);
            if (in.z != (uint32_t)fam_->z || in.rows < 1 || in.rows > (uint32_t)R || in.frames < 1 || in.frames > (uint32_t)nf ||
                in.first_frame > (uint32_t)nf - in.frames || in.workbits != (uint32_t)b.workbits ||
                !detail::payload_is(in.base, sizeof(in), 4ull * in.frames * in.rows * zw))
                return ERR_LDPC_BAD_PACKET;
            State &s = st_[b.startEpoch];
            if (s.frames == 0) {
                s.frames = nf;
                s.parity.assign((size_t)s.frames * R * zw, 0u);
                s.rows.assign(s.frames, 0);
                s.done.assign(s.frames, 0);
                s.was_revealed.assign(s.frames, 0);
                s.crc_sent.assign(s.frames, 0);
                s.received = 0;
            }
            for (uint32_t f = 0; f < in.frames; ++f)
                if (s.rows[in.first_frame + f] != 0) return ERR_LDPC_BAD_PACKET;   // a frame's first rows arrive once
            b.leakageBits += (int)(in.frames * in.rows) * fam_->z;
            for (uint32_t f = 0; f < in.frames; ++f) {
                std::memcpy(s.parity.data() + (size_t)(in.first_frame + f) * R * zw,
                            pkts[i] + sizeof(in) + (size_t)f * in.rows * zw * 4, (size_t)in.rows * zw * 4);
                s.rows[in.first_frame + f] = (int)in.rows;
            }
            s.received += (int)in.frames;
        }
        std::vector<Work> work;
        for (size_t i = 0; i < pkts.size(); ++i) {
            KeyBlock &b = *blocks[i];
            State &s = st_[b.startEpoch];
            if (s.received < s.frames || s.queued) continue;   // wait for the block's last parity packet
            s.queued = true;
            for (int f = 0; f < s.frames; ++f) work.push_back(Work{&b, f});
        }
        return decode_and_answer(work, send);
    }

    // handler for SUBTYPE_LDPC_MORE
    int on_more(const std::vector<KeyBlock *> &blocks, const std::vector<const char *> &pkts, std::vector<Packet> &send)
    {
        const int zw = fam_->zwords(), kw = fam_->kwords(), R = fam_->max_rows();
        std::vector<Work> work;
        for (size_t i = 0; i < pkts.size(); ++i) {
            EcPktHdr_LdpcMore in;
            if (!detail::read_header(pkts[i], SUBTYPE_LDPC_MORE, in)) return ERR_LDPC_BAD_PACKET;
            KeyBlock &b = *blocks[i];
            auto sit = st_.find(b.startEpoch);
            if (sit == st_.end()) return ERR_LDPC_NO_BLOCK;
            State &s = sit->second;
            if (in.n_frames < 1 || in.n_frames > (uint32_t)s.frames || in.row_to > (uint32_t)R || in.row_from > in.row_to ||
                (!in.reveal && in.row_from == in.row_to))
                return ERR_LDPC_BAD_PACKET;
            const size_t per = in.reveal ? (size_t)kw * 4 : (size_t)(in.row_to - in.row_from) * zw * 4;
            if (!detail::payload_is(in.base, sizeof(in), (4ull + per) * in.n_frames)) return ERR_LDPC_BAD_PACKET;
            const uint32_t *idx = reinterpret_cast<const uint32_t *>(pkts[i] + sizeof(in));
            for (uint32_t k = 0; k < in.n_frames; ++k)   // every index inside the block; new rows continue where the frame stands
                if (idx[k] >= (uint32_t)s.frames || (!in.reveal && (uint32_t)s.rows[idx[k]] != in.row_from)) return ERR_LDPC_BAD_PACKET;
            const char *payload = pkts[i] + sizeof(in) + (size_t)in.n_frames * 4;
            for (uint32_t k = 0; k < in.n_frames; ++k) {
                const int f = (int)idx[k];
                if (in.reveal) {   // Alice's key bits for this frame: copy them over
                    std::vector<uint32_t> mine(kw), hers(kw);
                    detail::copy_frame_key(b, f, kw, mine.data());
                    std::memcpy(hers.data(), payload + k * per, per);
                    store_frame(b, f, hers.data(), mine.data());
                    s.done[f] = 1;
                    ++s.revealed;
                    if (!s.was_revealed[f]) { b.leakageBits += fam_->K() - (s.crc_sent[f] ? 32 : 0); s.was_revealed[f] = 1; }
                    if (s.finished) { s.finished = false; ++s.pending; }   // a frame revealed after a CRC mismatch: answer again
                    touched_[&b] = 1;
                } else {
                    std::memcpy(s.parity.data() + ((size_t)f * R + in.row_from) * zw, payload + k * per, per);
                    s.rows[f] = (int)in.row_to;
                    b.leakageBits += (int)(in.row_to - in.row_from) * fam_->z;
                    work.push_back(Work{&b, f});
                }
            }
            s.pending -= (int)in.n_frames;
            if (s.pending < 0) s.pending = 0;
        }
        return decode_and_answer(work, send);
    }

    // the block's state is dropped when Alice has confirmed it (the caller learns that from her side of the protocol)
    void release(uint32_t epoch) { st_.erase(epoch); }
    // test hook: flip one key bit of the block right before the next DONE is built (exercises the CRC confirmation)
    void corrupt_before_done(uint32_t epoch) { corrupt_.push_back(epoch); }

    int rounds(uint32_t epoch) const { auto it = st_.find(epoch); return it == st_.end() ? 0 : it->second.round; }

private:
    struct State {
        int frames = 0, received = 0, round = 0, revealed = 0, pending = 0;
        bool queued = false, finished = false;
        std::vector<uint32_t> parity;
        std::vector<int> rows;
        std::vector<char> done, was_revealed, crc_sent;
    };
    struct Work { KeyBlock *b; int f; };

    void store_frame(KeyBlock &b, int f, const uint32_t *corrected, const uint32_t *old)
    {
        const int kw = fam_->kwords();
        const int total_words = (b.workbits + 31) / 32;
        for (int w = 0; w < kw; ++w) {
            const int dst = f * kw + w;
            if (dst >= total_words) break;
            uint32_t diff = corrected[w] ^ old[w];
            if (dst == total_words - 1 && (b.workbits & 31)) diff &= ~0u << (32 - (b.workbits & 31));
            b.correctedErrors += __builtin_popcount(diff);
            b.mainBufPtr[dst] ^= diff;
        }
    }

    // decode the listed frames (grouped by the number of rows they hold), then per block: NACK the failures or DONE
    int decode_and_answer(const std::vector<Work> &work, std::vector<Packet> &send)
    {
        const int zw = fam_->zwords(), kw = fam_->kwords(), R = fam_->max_rows();
        std::map<int, std::vector<Work>> by_rows;
        for (const Work &w : work) by_rows[st_.at(w.b->startEpoch).rows[w.f]].push_back(w);
        std::map<KeyBlock *, std::vector<uint32_t>> failed;
        std::map<KeyBlock *, char> touched;
        for (auto &kv : by_rows) {
            const int m = kv.first;
            const std::vector<Work> &ws = kv.second;
            qldpc_decoder *dec;
            try {
                dec = fam_->decoder(m);
            } catch (const std::exception &) {
                return ERR_LDPC_UNSUPPORTED;
            }
            const int cw = qldpc_codeword_words(dec), F = (int)ws.size();
            std::vector<uint32_t> bits((size_t)F * cw), out((size_t)F * kw), known(cw, 0u);
            for (int w = kw; w < cw; ++w) known[w] = ~0u;   // every parity bit of the rate-matched code was received
            std::vector<uint8_t> ok(F);
            // frames of one rate share a QBER only if they come from one block: decode block by block when they differ
            int a = 0;
            while (a < F) {
                int e = a;
                while (e < F && ws[e].b->localError == ws[a].b->localError) ++e;
                for (int k = a; k < e; ++k) {
                    detail::copy_frame_key(*ws[k].b, ws[k].f, kw, bits.data() + (size_t)k * cw);
                    std::memcpy(bits.data() + (size_t)k * cw + kw,
                                st_.at(ws[k].b->startEpoch).parity.data() + (size_t)ws[k].f * R * zw, (size_t)m * zw * 4);
                }
                const float q = std::min(0.45f, std::max(1e-4f, ws[a].b->localError));
                const float llr = std::min(30.0f, 4.0f * std::log((1.0f - q) / q));   // int8 scale 2^2
                const int rc = qldpc_decode_bits(dec, bits.data() + (size_t)a * cw, known.data(), nullptr, std::round(llr), 31.0f,
                                                 nullptr, e - a, out.data() + (size_t)a * kw, ok.data() + a, nullptr);
                if (rc) return ERR_LDPC_UNSUPPORTED;
                a = e;
            }
            for (int k = 0; k < F; ++k) {
                KeyBlock *b = ws[k].b;
                State &s = st_.at(b->startEpoch);
                touched[b] = 1;
                if (ok[k]) {
                    store_frame(*b, ws[k].f, out.data() + (size_t)k * kw, bits.data() + (size_t)k * cw);
                    s.done[ws[k].f] = 1;
                } else {
                    failed[b].push_back((uint32_t)ws[k].f);
                }
            }
        }
        for (auto &kv : touched) {
            KeyBlock *b = kv.first;
            State &s = st_.at(b->startEpoch);
            auto it = failed.find(b);
            if (it != failed.end() && !it->second.empty()) {
                ++s.round;
                s.pending += (int)it->second.size();
                EcPktHdr_LdpcNack h{};
                h.round = (uint32_t)s.round; h.n_failed = (uint32_t)it->second.size();
                Packet p = detail::make_packet(SUBTYPE_LDPC_NACK, *b, h, it->second.size() * 4);
                std::memcpy(p.data() + sizeof(h), it->second.data(), it->second.size() * 4);
                send.push_back(std::move(p));
            }
        }
        // a block is done when every frame is (also reached through on_more's reveal path): DONE carries one CRC per frame
        for (auto &kv : touched_) touched[kv.first] = 1;
        touched_.clear();
        for (auto &kv : touched) {
            KeyBlock *b = kv.first;
            auto it = st_.find(b->startEpoch);
            if (it == st_.end()) continue;
            State &s = it->second;
            bool all = s.frames > 0 && s.pending == 0 && !s.finished;
            for (char d : s.done) all = all && d;
            if (!all) continue;
            for (auto c = corrupt_.begin(); c != corrupt_.end(); ++c)
                if (*c == b->startEpoch) { b->mainBufPtr[0] ^= 0x00010000u; corrupt_.erase(c); break; }
            std::vector<uint32_t> crc;
            if (detail::block_crcs(*b, fam_->K(), fam_->kwords(), fam_->prm.device, crc)) return ERR_LDPC_UNSUPPORTED;
            for (int f = 0; f < s.frames; ++f)   // 32 disclosed parities per frame, once (Alice counts the same on receipt)
                if (!s.crc_sent[f]) { s.crc_sent[f] = 1; if (!s.was_revealed[f]) b->leakageBits += 32; }
            EcPktHdr_LdpcDone h{};
            h.rounds = (uint32_t)s.round; h.frames_revealed = (uint32_t)s.revealed; h.frames = (uint32_t)crc.size();
            h.corrected_errors = (uint32_t)b->correctedErrors;
            Packet p = detail::make_packet(SUBTYPE_LDPC_DONE, *b, h, crc.size() * 4);
            std::memcpy(p.data() + sizeof(h), crc.data(), crc.size() * 4);
            send.push_back(std::move(p));
            last_rounds_[b->startEpoch] = s.round;
            s.finished = true;
        }
        return 0;
    }

public:
    std::map<uint32_t, int> last_rounds_;   // rounds a finished block needed (notify-pipe statistics)

private:
    std::shared_ptr<CodeFamily> fam_;
    std::map<uint32_t, State> st_;
    std::map<KeyBlock *, char> touched_;   // blocks changed by a reveal, to be answered by the next decode_and_answer
    std::vector<uint32_t> corrupt_;
};

}  // namespace ecd2
}  // namespace qldpc
