// C face of the blind-reconciliation handlers (include/qldpc_ecd2.h) over host/qldpc_blind.hpp: what the ecd2 packet
// handlers of integration/ecd2_ldpc.patch call.  Exceptions stop here; ecd2 error codes go out.
#include <new>

#include "../../include/qldpc_ecd2.h"
#include "../host/qldpc_blind.hpp"

using namespace qldpc::ecd2;

struct qldpc_ecd2 {
    std::shared_ptr<CodeFamily> fam;
    std::unique_ptr<BlindAlice> alice;
    std::unique_ptr<BlindBob> bob;
    std::vector<Packet> out;
};

namespace {
KeyBlock to_key_block(const qldpc_ecd2_block &b)
{
    KeyBlock k;
    k.startEpoch = b.start_epoch;
    k.numberOfEpochs = b.number_of_epochs;
    k.mainBufPtr = b.main_buf;
    k.workbits = b.workbits;
    k.localError = b.local_error;
    k.leakageBits = b.leakage_bits;
    k.correctedErrors = b.corrected_errors;
    return k;
}
void from_key_block(const KeyBlock &k, qldpc_ecd2_block &b)
{
    b.leakage_bits = k.leakageBits;
    b.corrected_errors = k.correctedErrors;
}
}  // namespace

extern "C" void qldpc_ecd2_config_default(qldpc_ecd2_config *cfg)
{
    if (!cfg) return;
    const Params p;
    cfg->base_qc = nullptr;
    cfg->device = p.device;
    cfg->f_start = p.f_start;
    cfg->delta_rows = p.delta_rows;
    cfg->max_iter = p.max_iter;
    cfg->frames_per_packet = p.frames_per_packet;
}

extern "C" int qldpc_ecd2_open(const qldpc_ecd2_config *cfg, qldpc_ecd2 **out)
{
    if (!cfg || !out || !cfg->base_qc) return ERR_LDPC_UNSUPPORTED;
    try {
        Params p;
        p.base_qc = cfg->base_qc;
        p.device = cfg->device;
        if (cfg->f_start > 0) p.f_start = cfg->f_start;
        if (cfg->delta_rows > 0) p.delta_rows = cfg->delta_rows;
        if (cfg->max_iter > 0) p.max_iter = cfg->max_iter;
        if (cfg->frames_per_packet > 0) p.frames_per_packet = cfg->frames_per_packet;
        std::unique_ptr<qldpc_ecd2> ctx(new qldpc_ecd2());
        ctx->fam = std::make_shared<CodeFamily>(p);
        ctx->fam->decoder(ctx->fam->max_rows());   // fails here without an sm_100 device: no CPU fallback
        ctx->alice.reset(new BlindAlice(ctx->fam));
        ctx->bob.reset(new BlindBob(ctx->fam));
        *out = ctx.release();
        return 0;
    } catch (const std::exception &) {
        return ERR_LDPC_UNSUPPORTED;
    }
}

extern "C" void qldpc_ecd2_close(qldpc_ecd2 *ctx) { delete ctx; }

extern "C" int qldpc_ecd2_initiate(qldpc_ecd2 *ctx, qldpc_ecd2_block *blk)
{
    if (!ctx || !blk || !blk->main_buf || blk->workbits <= 0) return ERR_LDPC_UNSUPPORTED;
    try {
        ctx->out.clear();
        KeyBlock k = to_key_block(*blk);
        std::vector<KeyBlock *> blocks{&k};
        const int rc = ctx->alice->initiate(blocks, ctx->out);
        from_key_block(k, *blk);
        return rc;
    } catch (const std::exception &) {
        return ERR_LDPC_UNSUPPORTED;
    }
}

extern "C" int qldpc_ecd2_handle(qldpc_ecd2 *ctx, qldpc_ecd2_block *blk, const char *receivebuf, int *finished)
{
    if (!ctx || !blk || !receivebuf || !blk->main_buf) return ERR_LDPC_UNSUPPORTED;
    if (finished) *finished = 0;
    try {
        ctx->out.clear();
        EcPktHdr_Base h;
        std::memcpy(&h, receivebuf, sizeof(h));
        KeyBlock k = to_key_block(*blk);
        std::vector<KeyBlock *> blocks{&k};
        std::vector<const char *> pkts{receivebuf};
        int rc;
        switch (h.subtype) {
        case SUBTYPE_LDPC_PARITY: rc = ctx->bob->on_parity(blocks, pkts, ctx->out); break;
        case SUBTYPE_LDPC_MORE: rc = ctx->bob->on_more(blocks, pkts, ctx->out); break;
        case SUBTYPE_LDPC_NACK: rc = ctx->alice->on_nack(k, receivebuf, ctx->out); break;
        case SUBTYPE_LDPC_DONE: {
            bool confirmed = false;
            rc = ctx->alice->on_done(k, receivebuf, ctx->out, confirmed);
            if (finished && confirmed) *finished = 1;
            break;
        }
        default: rc = 45; break;   // errormessage[45] "received unrecognized message subtype"
        }
        from_key_block(k, *blk);
        return rc;
    } catch (const std::exception &) {
        return ERR_LDPC_UNSUPPORTED;
    }
}

extern "C" void qldpc_ecd2_release(qldpc_ecd2 *ctx, uint32_t start_epoch)
{
    if (ctx) ctx->bob->release(start_epoch);
}

extern "C" int32_t qldpc_ecd2_packet_count(const qldpc_ecd2 *ctx) { return ctx ? (int32_t)ctx->out.size() : 0; }

extern "C" const char *qldpc_ecd2_packet_data(const qldpc_ecd2 *ctx, int32_t index, uint32_t *length_in_bytes)
{
    if (!ctx || index < 0 || index >= (int32_t)ctx->out.size()) return nullptr;
    if (length_in_bytes) *length_in_bytes = (uint32_t)ctx->out[index].size();
    return reinterpret_cast<const char *>(ctx->out[index].data());
}
