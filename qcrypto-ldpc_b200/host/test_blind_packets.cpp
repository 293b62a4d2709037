// Malformed-packet tests of the ecd2 LDPC handlers (qldpc_blind.hpp): every received field is validated before use.
//   test_blind_packets <base.qc> bob     Bob's handlers, no GPU needed (rejection happens before any decode)
//   test_blind_packets <base.qc> alice   Alice's handlers (her state needs one encode: GPU)
#include <cstdio>

#include "qldpc_blind.hpp"

using namespace qldpc::ecd2;

static int failures = 0;
#define EXPECT(cond) do { if (!(cond)) { std::printf("FAILED line %d: %s\n", __LINE__, #cond); ++failures; } } while (0)

template <class H>
static Packet pkt(uint32_t subtype, uint32_t epoch, H h, size_t payload, int len_delta = 0)
{
    Packet p(sizeof(H) + payload, 0);
    h.base.tag = EC_PACKET_TAG;
    h.base.subtype = subtype;
    h.base.epoch = epoch;
    h.base.numberOfEpochs = 1;
    h.base.totalLengthInBytes = (uint32_t)((long)p.size() + len_delta);
    std::memcpy(p.data(), &h, sizeof(H));
    return p;
}

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    Params prm;
    prm.base_qc = argv[1];
    auto fam = std::make_shared<CodeFamily>(prm);
    const int zw = fam->zwords(), R = fam->max_rows(), kw = fam->kwords();
    const int workbits = 3 * fam->K() - 100;   // three frames
    std::vector<uint32_t> key((workbits + 31) / 32, 0x12345678u);
    KeyBlock blk;
    blk.startEpoch = 77;
    blk.mainBufPtr = key.data();
    blk.workbits = workbits;
    blk.localError = 0.03f;
    std::vector<Packet> send;

    if (std::string(argv[2]) == "bob") {
        BlindBob bob(fam);
        auto parity = [&](uint32_t z, uint32_t frames, uint32_t first, uint32_t rows, uint32_t wb, int len_delta) {
            EcPktHdr_LdpcParity h{};
            h.z = z; h.frames = frames; h.first_frame = first; h.rows = rows; h.workbits = wb; h.qber = 0.03f;
            const size_t pay = (rows <= 64u && frames <= 16u) ? (size_t)frames * rows * zw * 4 : 64;
            Packet p = pkt(SUBTYPE_LDPC_PARITY, 77, h, pay, len_delta);
            std::vector<KeyBlock *> b{&blk};
            std::vector<const char *> pk{(const char *)p.data()};
            return bob.on_parity(b, pk, send);
        };
        EXPECT(parity(fam->z + 32, 1, 0, 6, workbits, 0) == ERR_LDPC_BAD_PACKET);   // wrong lifting size
        EXPECT(parity(fam->z, 1, 0, R + 1, workbits, 0) == ERR_LDPC_BAD_PACKET);     // more rows than the base graph has
        EXPECT(parity(fam->z, 1, 0, 0, workbits, 0) == ERR_LDPC_BAD_PACKET);
        EXPECT(parity(fam->z, 4, 0, 6, workbits, 0) == ERR_LDPC_BAD_PACKET);         // more frames than the block has
        EXPECT(parity(fam->z, 2, 2, 6, workbits, 0) == ERR_LDPC_BAD_PACKET);         // first_frame + frames beyond the block
        EXPECT(parity(fam->z, 1, 0xffffffffu, 6, workbits, 0) == ERR_LDPC_BAD_PACKET);
        EXPECT(parity(fam->z, 1, 0, 6, workbits + 1, 0) == ERR_LDPC_BAD_PACKET);     // block length disagrees
        EXPECT(parity(fam->z, 1, 0, 6, workbits, -4) == ERR_LDPC_BAD_PACKET);        // declared length shorter than the payload
        EXPECT(parity(fam->z, 1, 0, 6, workbits, 8) == ERR_LDPC_BAD_PACKET);
        // MORE for a block Bob has never seen
        {
            EcPktHdr_LdpcMore h{};
            h.n_frames = 1; h.row_from = 6; h.row_to = 8;
            Packet p = pkt(SUBTYPE_LDPC_MORE, 77, h, 4 + 2 * zw * 4);
            std::vector<KeyBlock *> b{&blk};
            std::vector<const char *> pk{(const char *)p.data()};
            EXPECT(bob.on_more(b, pk, send) == ERR_LDPC_NO_BLOCK);
        }
        // a well-formed first packet creates the state (1 of 3 frames: nothing is decoded yet, so no GPU is needed) ...
        EXPECT(parity(fam->z, 1, 0, 6, workbits, 0) == 0);
        EXPECT(blk.leakageBits == 6 * fam->z);
        EXPECT(parity(fam->z, 1, 0, 6, workbits, 0) == ERR_LDPC_BAD_PACKET);         // the same frame's first rows twice
        auto more = [&](uint32_t n, uint32_t idx0, uint32_t from, uint32_t to, uint32_t reveal, int len_delta) {
            EcPktHdr_LdpcMore h{};
            h.n_frames = n; h.row_from = from; h.row_to = to; h.reveal = reveal;
            const size_t per = reveal ? (size_t)kw * 4 : (to >= from && to - from <= 64u ? (size_t)(to - from) * zw * 4 : 0);
            Packet p = pkt(SUBTYPE_LDPC_MORE, 77, h, (n <= 16u ? n : 1) * (4 + per), len_delta);
            std::memcpy(p.data() + sizeof(h), &idx0, 4);
            std::vector<KeyBlock *> b{&blk};
            std::vector<const char *> pk{(const char *)p.data()};
            return bob.on_more(b, pk, send);
        };
        // ... against which MORE packets are checked
        EXPECT(more(1, 3, 6, 8, 0, 0) == ERR_LDPC_BAD_PACKET);            // frame index outside the block
        EXPECT(more(1, 0xfffffff0u, 6, 8, 0, 0) == ERR_LDPC_BAD_PACKET);
        EXPECT(more(1, 0, 6, R + 1, 0, 0) == ERR_LDPC_BAD_PACKET);        // row_to beyond the base graph
        EXPECT(more(1, 0, 8, 6, 0, 0) == ERR_LDPC_BAD_PACKET);            // row_from > row_to
        EXPECT(more(1, 0, 4, 6, 0, 0) == ERR_LDPC_BAD_PACKET);            // rows do not continue where the frame stands (6)
        EXPECT(more(1, 0, 6, 6, 0, 0) == ERR_LDPC_BAD_PACKET);            // empty row range
        EXPECT(more(4, 0, 6, 8, 0, 0) == ERR_LDPC_BAD_PACKET);            // more frames than the block has
        EXPECT(more(1, 0, 6, 8, 0, -4) == ERR_LDPC_BAD_PACKET);
        EXPECT(more(1, 0, R, R, 1, 4) == ERR_LDPC_BAD_PACKET);
        {   // wrong tag / subtype
            EcPktHdr_LdpcMore h{};
            h.n_frames = 1; h.row_from = 6; h.row_to = 8;
            Packet p = pkt(SUBTYPE_LDPC_NACK, 77, h, 4 + 2 * zw * 4);
            std::vector<KeyBlock *> b{&blk};
            std::vector<const char *> pk{(const char *)p.data()};
            EXPECT(bob.on_more(b, pk, send) == ERR_LDPC_BAD_PACKET);
        }
        EXPECT(send.empty());
        if (!failures) std::printf("bob: all malformed packets rejected\n");
        return failures ? 1 : 0;
    }

    BlindAlice alice(fam);
    std::vector<KeyBlock *> blocks{&blk};
    if (int rc = alice.initiate(blocks, send)) { std::printf("initiate failed: %d\n", rc); return 3; }
    const int leak0 = blk.leakageBits;
    send.clear();
    auto nack = [&](uint32_t n, uint32_t idx0, int len_delta, uint32_t epoch = 77) {
        EcPktHdr_LdpcNack h{};
        h.round = 1; h.n_failed = n;
        Packet p = pkt(SUBTYPE_LDPC_NACK, epoch, h, (n <= 16u ? n : 1) * 4, len_delta);
        std::memcpy(p.data() + sizeof(h), &idx0, 4);
        KeyBlock other = blk;
        other.startEpoch = epoch;
        return alice.on_nack(epoch == 77 ? blk : other, (const char *)p.data(), send);
    };
    EXPECT(nack(1, 3, 0) == ERR_LDPC_BAD_PACKET);             // frame index outside the block
    EXPECT(nack(1, 0x80000000u, 0) == ERR_LDPC_BAD_PACKET);
    EXPECT(nack(0, 0, 0) == ERR_LDPC_BAD_PACKET);
    EXPECT(nack(4, 0, 0) == ERR_LDPC_BAD_PACKET);             // more failures than frames
    EXPECT(nack(1, 0, -4) == ERR_LDPC_BAD_PACKET);            // declared length shorter than the index list
    EXPECT(nack(1, 0, 0, 78) == ERR_LDPC_NO_BLOCK);
    EXPECT(send.empty() && blk.leakageBits == leak0);
    {
        EcPktHdr_LdpcDone h{};
        h.frames = 5;                                          // the block has three frames
        Packet p = pkt(SUBTYPE_LDPC_DONE, 77, h, 5 * 4);
        bool confirmed = true;
        EXPECT(alice.on_done(blk, (const char *)p.data(), send, confirmed) == ERR_LDPC_BAD_PACKET);
        h.frames = 3;
        p = pkt(SUBTYPE_LDPC_DONE, 77, h, 3 * 4, -4);
        EXPECT(alice.on_done(blk, (const char *)p.data(), send, confirmed) == ERR_LDPC_BAD_PACKET);
        EXPECT(blk.leakageBits == leak0);
    }
    EXPECT(nack(1, 1, 0) == 0 && send.size() == 1);           // a well-formed NACK is answered
    if (!failures) std::printf("alice: all malformed packets rejected\n");
    return failures ? 1 : 0;
}
