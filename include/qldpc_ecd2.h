/*
 * qldpc_ecd2.h -- C face of the blind LDPC reconciliation plug-in for the ecd2 daemon (libqldpc_b200.so).
 *
 * What calls this: the packet handlers that integration/ecd2_ldpc.patch adds to the reference tree
 * (errorcorrection/subcomponents/ldpc_reconcile.c), registered as PacketHandlerArray entries
 * (errorcorrection/definitions/algorithms/packet_manager.h:33) for the two algorithm slots the reference reserves and
 * answers with error 81 today: ALG_LDPC_CONTINUE_ROLES / ALG_LDPC_FLIP_ROLES
 * (errorcorrection/definitions/algorithms/algorithms.h:36-42, errorcorrection/subcomponents/qber_estim.c:337-340,420-423).
 * The protocol, packet layouts (EC subtypes 9..12 after errorcorrection/definitions/packets.h:46-56) and leakage
 * accounting are those of qcrypto-ldpc_b200/host/qldpc_blind.hpp.
 *
 * Plain C, no exceptions: every call returns an ecd2 error code (0 = ok, else an index into errormessage[],
 * errorcorrection/ecd2.h:251-337; 81 "Unsupported functionality" when no sm_100 device / library problem,
 * 49 unknown block, 85 malformed packet).  One context per daemon; ecd2 is single-threaded
 * (errorcorrection/ecd2.c:481) and so is this interface.
 */
#ifndef QLDPC_ECD2_H
#define QLDPC_ECD2_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qldpc_ecd2 qldpc_ecd2;

/* the ProcessBlock fields an LDPC handler reads and writes (errorcorrection/definitions/processblock.h:102-129);
 * the handler copies them in before a call and back out after it */
typedef struct qldpc_ecd2_block {
    uint32_t start_epoch, number_of_epochs;
    uint32_t *main_buf;        /* pb->mainBufPtr: key bits, MSB-first words (helpers.h:65-68); corrected in place (follower) */
    int32_t workbits;          /* pb->workbits after helper_cleanupRevealedBits (helpers.c:31-69) */
    float local_error;         /* pb->localError: estimated QBER */
    int32_t leakage_bits;      /* pb->leakageBits: += every parity / revealed / CRC bit (priv_amp.c:47,166) */
    int32_t corrected_errors;  /* pb->correctedErrors */
} qldpc_ecd2_block;

typedef struct qldpc_ecd2_config {
    const char *base_qc;       /* NR base graph (.qc), e.g. ldpc_examples/.../matrices/H/NR_1_1_192.qc or an NR_*_384 table */
    int32_t device;            /* CUDA ordinal */
    float f_start;             /* initial efficiency target (first parity rows = ceil(f_start * 22 * h(QBER))), default 1.25 */
    int32_t delta_rows;        /* parity block rows added per NACK round, default 2 */
    int32_t max_iter;          /* default 20 */
    int32_t frames_per_packet; /* default 4: below transferd's 10 000-byte EC packet cap (remotecrypto/transferd.h:139) */
} qldpc_ecd2_config;

void qldpc_ecd2_config_default(qldpc_ecd2_config *cfg);
/* fails with 81 when the base graph cannot be read or no sm_100 device is present (there is no CPU decoder behind this) */
int  qldpc_ecd2_open(const qldpc_ecd2_config *cfg, qldpc_ecd2 **out);
void qldpc_ecd2_close(qldpc_ecd2 *ctx);

/* EC initiator, body of `case ALG_LDPC_*` in qber_prepareErrorCorrection (qber_estim.c:420-423): NR-encode the block's frames,
 * queue the first parity rows (subtype 9).  Outgoing packets are collected in the context, see qldpc_ecd2_packet_*. */
int  qldpc_ecd2_initiate(qldpc_ecd2 *ctx, qldpc_ecd2_block *blk);
/* one received packet of subtype 9..12 (`receivebuf` of a PacketHandlerArray entry).  *finished is set to 1 when the
 * initiator has confirmed the block (all CRCs equal): the caller then runs privAmp_sendPrivAmpMsgAndPrivAmp
 * (cascade_biconf.c:892).  The follower learns the end from the privacy-amplification packet (subtype 8). */
int  qldpc_ecd2_handle(qldpc_ecd2 *ctx, qldpc_ecd2_block *blk, const char *receivebuf, int *finished);
/* drop the per-block protocol state (freeData of the algorithm data manager, data_manager.h:28-32) */
void qldpc_ecd2_release(qldpc_ecd2 *ctx, uint32_t start_epoch);

/* packets produced by the last initiate / handle call; the caller copies each into a malloc2'd buffer and hands it to
 * comms_insertSendPacket (comms.c:16-38), which takes ownership of that copy */
int32_t     qldpc_ecd2_packet_count(const qldpc_ecd2 *ctx);
const char *qldpc_ecd2_packet_data(const qldpc_ecd2 *ctx, int32_t index, uint32_t *length_in_bytes);

#ifdef __cplusplus
}
#endif
#endif /* QLDPC_ECD2_H */
