// Two-party loop-back of the ecd2 LDPC plug-in (qldpc_blind.hpp): what two ecd2 daemons joined by their send/receive
// FIFOs do for the LDPC algorithm slots, minus the daemon around it (EC/ecd2.c:407-550 reads a packet, looks up the
// process block by epoch and calls the handler of its subtype, :525-526; outgoing packets are queued, :422-427).
//
//   driver_blind <base.qc> <keys.bin> <corrected.bin> [f_start] [delta_rows] [max_iter]
// keys.bin     : int32 n_blocks, int32 workbits, float qber, then per block ceil(workbits/32) words of Alice's key and
//                as many of Bob's (MSB-first words, the stream-3 payload layout, packetheaders/pkt_header_3.h:7-12)
// corrected.bin: Bob's blocks after reconciliation, same layout (one copy)
// stdout       : one JSON line (rounds, leakage, efficiency, packets, wall time, key bits per second)
// exit codes   : 0 ok, 2 usage, 3 decoder unavailable (no sm_100 device), 4 keys differ after reconciliation
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "qldpc_blind.hpp"

using namespace qldpc::ecd2;

int main(int argc, char **argv)
{
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s <base.qc> <keys.bin> <corrected.bin> [f_start] [delta_rows] [max_iter]\n", argv[0]);
        return 2;
    }
    Params prm;
    prm.base_qc = argv[1];
    if (argc > 4) prm.f_start = (float)std::atof(argv[4]);
    if (argc > 5) prm.delta_rows = std::atoi(argv[5]);
    if (argc > 6) prm.max_iter = std::atoi(argv[6]);
    FILE *in = std::fopen(argv[2], "rb");
    if (!in) { std::perror(argv[2]); return 2; }
    int32_t n_blocks = 0, workbits = 0;
    float qber = 0;
    if (std::fread(&n_blocks, 4, 1, in) != 1 || std::fread(&workbits, 4, 1, in) != 1 || std::fread(&qber, 4, 1, in) != 1) return 2;
    const int words = (workbits + 31) / 32;
    std::vector<std::vector<uint32_t>> ka(n_blocks, std::vector<uint32_t>(words)), kb = ka;
    for (int b = 0; b < n_blocks; ++b)
        if (std::fread(ka[b].data(), 4, words, in) != (size_t)words || std::fread(kb[b].data(), 4, words, in) != (size_t)words) return 2;
    std::fclose(in);

    std::vector<KeyBlock> A(n_blocks), B(n_blocks);
    std::vector<KeyBlock *> pa, pb;
    for (int b = 0; b < n_blocks; ++b) {
        A[b].startEpoch = B[b].startEpoch = 0xb0b80000u + (uint32_t)b;
        A[b].mainBufPtr = ka[b].data(); B[b].mainBufPtr = kb[b].data();
        A[b].workbits = B[b].workbits = workbits;
        A[b].localError = B[b].localError = qber;
        pa.push_back(&A[b]); pb.push_back(&B[b]);
    }
    auto block_of = [&](std::vector<KeyBlock> &v, uint32_t epoch) -> KeyBlock * { return &v[epoch - 0xb0b80000u]; };

    try {
        auto fam = std::make_shared<CodeFamily>(prm);
        fam->decoder(fam->max_rows());   // fails here when there is no device: no CPU fallback
        BlindAlice alice(fam);
        BlindBob bob(fam);
        // warm-up: create the decoders of the rates the run will touch (cuModuleLoad, table upload) outside the timing
        for (int m = fam->initial_rows(qber); m <= fam->max_rows(); m += prm.delta_rows) fam->decoder(m);

        size_t bytes_ab = 0, bytes_ba = 0, pk_ab = 0, pk_ba = 0;
        int done = 0, turns = 0;
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<Packet> a2b, b2a;
        if (std::getenv("QLDPC_BLIND_CORRUPT")) bob.corrupt_before_done(B[0].startEpoch);   // test: provoke a CRC mismatch
        if (int rc = alice.initiate(pa, a2b)) { std::fprintf(stderr, "initiate: error %d\n", rc); return 3; }
        while (done < n_blocks && turns < 1000) {
            ++turns;
            // Bob's select loop: everything pending on the receive FIFO is dispatched by subtype
            std::vector<KeyBlock *> blk9, blk11;
            std::vector<const char *> p9, p11;
            for (auto &p : a2b) {
                EcPktHdr_Base h;
                std::memcpy(&h, p.data(), sizeof(h));
                bytes_ab += p.size(); ++pk_ab;
                if (h.subtype == SUBTYPE_LDPC_PARITY) { blk9.push_back(block_of(B, h.epoch)); p9.push_back((const char *)p.data()); }
                else if (h.subtype == SUBTYPE_LDPC_MORE) { blk11.push_back(block_of(B, h.epoch)); p11.push_back((const char *)p.data()); }
            }
            b2a.clear();
            if (!p9.empty())
                if (int rc = bob.on_parity(blk9, p9, b2a)) { std::fprintf(stderr, "on_parity: error %d\n", rc); return 3; }
            if (!p11.empty())
                if (int rc = bob.on_more(blk11, p11, b2a)) { std::fprintf(stderr, "on_more: error %d\n", rc); return 3; }
            a2b.clear();
            // Alice's select loop
            for (auto &p : b2a) {
                EcPktHdr_Base h;
                std::memcpy(&h, p.data(), sizeof(h));
                bytes_ba += p.size(); ++pk_ba;
                KeyBlock *blk = block_of(A, h.epoch);
                if (h.subtype == SUBTYPE_LDPC_NACK) {
                    if (int rc = alice.on_nack(*blk, (const char *)p.data(), a2b)) { std::fprintf(stderr, "on_nack: error %d\n", rc); return 3; }
                }
                else if (h.subtype == SUBTYPE_LDPC_DONE) {
                    bool confirmed = false;
                    if (int rc = alice.on_done(*blk, (const char *)p.data(), a2b, confirmed)) { std::fprintf(stderr, "on_done: error %d\n", rc); return 3; }
                    if (confirmed) { bob.release(h.epoch); ++done; }
                }
            }
        }
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

        long leak = 0, leak_bob = 0, corrected = 0, diff_blocks = 0;
        std::map<int, int> round_hist;
        for (int b = 0; b < n_blocks; ++b) {
            leak += A[b].leakageBits;
            leak_bob += B[b].leakageBits;
            corrected += B[b].correctedErrors;
            if (ka[b] != kb[b]) ++diff_blocks;
            round_hist[bob.last_rounds_[A[b].startEpoch]]++;
        }
        FILE *out = std::fopen(argv[3], "wb");
        if (out) {
            for (int b = 0; b < n_blocks; ++b) std::fwrite(kb[b].data(), 4, words, out);
            std::fclose(out);
        }
        const double key_bits = (double)n_blocks * workbits;
        std::printf("{\"blocks\": %d, \"workbits\": %d, \"qber\": %.4f, \"frames_per_block\": %d, \"initial_rows\": %d, "
                    "\"turns\": %d, \"done\": %d, \"blocks_differ\": %ld, \"leak_bits\": %ld, \"leak_bits_bob\": %ld, \"corrected_errors\": %ld, "
                    "\"efficiency\": %.4f, \"packets_a2b\": %zu, \"packets_b2a\": %zu, \"bytes_a2b\": %zu, \"bytes_b2a\": %zu, "
                    "\"wall_s\": %.6f, \"reconciled_key_bits_per_s\": %.1f, \"crc_mismatches\": %d, \"round_hist\": {",
                    n_blocks, workbits, qber, (workbits + fam->K() - 1) / fam->K(), fam->initial_rows(qber), turns, done,
                    diff_blocks, leak, leak_bob, corrected, leak / (key_bits * h2(qber)), pk_ab, pk_ba, bytes_ab, bytes_ba, wall,
                    key_bits / wall, alice.mismatches());
        bool first = true;
        for (auto &kv : round_hist) { std::printf("%s\"%d\": %d", first ? "" : ", ", kv.first, kv.second); first = false; }
        std::printf("}}\n");
        return diff_blocks ? 4 : 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 3;
    }
}
