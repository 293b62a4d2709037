/**
 * @file ldpc_reconcile.c
 * @brief ecd2 packet handlers of the blind LDPC reconciliation; the decoding runs on a B200 behind qldpc_ecd2.h.
 *
 * Every handler copies the ProcessBlock fields the protocol works on into a qldpc_ecd2_block, makes ONE call into the
 * library, copies leakageBits / correctedErrors back and queues the packets the library produced with
 * comms_insertSendPacket (which owns the malloc2'd copies from then on, comms.c:16-38).
 */
#include "ldpc_reconcile.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "comms.h"
#include "helpers.h"
#include "priv_amp.h"
#include "qldpc_ecd2.h"

static qldpc_ecd2 *ldpcContext = NULL;   ///< one decoder context per daemon, opened with the first block
static int ldpcContextError = 0;         ///< sticky: a daemon without a B200 answers 81 for every block

static int ldpc_getContext(qldpc_ecd2 **ctx) {
  if (ldpcContext) { *ctx = ldpcContext; return 0; }
  if (ldpcContextError) return ldpcContextError;
  qldpc_ecd2_config cfg;
  qldpc_ecd2_config_default(&cfg);
  cfg.base_qc = getenv("ECD2_LDPC_BASE_QC");
  if (getenv("ECD2_LDPC_DEVICE")) cfg.device = atoi(getenv("ECD2_LDPC_DEVICE"));
  if (getenv("ECD2_LDPC_F_START")) cfg.f_start = (float)atof(getenv("ECD2_LDPC_F_START"));
  if (getenv("ECD2_LDPC_DELTA_ROWS")) cfg.delta_rows = atoi(getenv("ECD2_LDPC_DELTA_ROWS"));
  if (getenv("ECD2_LDPC_MAX_ITER")) cfg.max_iter = atoi(getenv("ECD2_LDPC_MAX_ITER"));
  ldpcContextError = qldpc_ecd2_open(&cfg, &ldpcContext);
  if (ldpcContextError) {
    fprintf(stderr, "LDPC reconciliation unavailable (no sm_100 device, or ECD2_LDPC_BASE_QC unset/unreadable): error %d\n",
            ldpcContextError);
    ldpcContext = NULL;
    return ldpcContextError;
  }
  *ctx = ldpcContext;
  return 0;
}

static void ldpc_blockFromProcessBlock(ProcessBlock *pb, qldpc_ecd2_block *blk) {
  blk->start_epoch = pb->startEpoch;
  blk->number_of_epochs = pb->numberOfEpochs;
  blk->main_buf = pb->mainBufPtr;
  blk->workbits = pb->workbits;
  blk->local_error = pb->localError;
  blk->leakage_bits = pb->leakageBits;
  blk->corrected_errors = pb->correctedErrors;
}

/// queue what the library produced; the send queue frees the copies after the write (ecd2.c:181-186)
static int ldpc_sendProducedPackets(qldpc_ecd2 *ctx) {
  int i, errorCode;
  for (i = 0; i < qldpc_ecd2_packet_count(ctx); i++) {
    unsigned int length = 0;
    const char *data = qldpc_ecd2_packet_data(ctx, i, &length);
    char *copy = (char *)malloc2(length);
    if (!copy) return 43;
    memcpy(copy, data, length);
    errorCode = comms_insertSendPacket(copy, length);
    if (errorCode) return errorCode;
  }
  return 0;
}

void ldpc_prepareBlock(ProcessBlock *pb) {
  helper_cleanupRevealedBits(pb);   /* sets pb->workbits, resets pb->leakageBits */
  pb->correctedErrors = 0;
}

int ldpc_initiateAfterQber(ProcessBlock *pb) {
  qldpc_ecd2 *ctx = NULL;
  qldpc_ecd2_block blk;
  int errorCode = ldpc_getContext(&ctx);
  if (errorCode) return errorCode;
  ldpc_prepareBlock(pb);
  ldpc_blockFromProcessBlock(pb, &blk);
  errorCode = qldpc_ecd2_initiate(ctx, &blk);
  if (errorCode) return errorCode;
  pb->leakageBits = blk.leakage_bits;
  return ldpc_sendProducedPackets(ctx);
}

/// common body of the four handlers
static int ldpc_handle(ProcessBlock *pb, char *receivebuf, PROCESSOR_ROLE expectedRole) {
  qldpc_ecd2 *ctx = NULL;
  qldpc_ecd2_block blk;
  int finished = 0;
  int errorCode;
  if (pb->processorRole != expectedRole) return 45;
  errorCode = ldpc_getContext(&ctx);
  if (errorCode) return errorCode;
  ldpc_blockFromProcessBlock(pb, &blk);
  errorCode = qldpc_ecd2_handle(ctx, &blk, receivebuf, &finished);
  pb->leakageBits = blk.leakage_bits;
  pb->correctedErrors = blk.corrected_errors;
  if (errorCode) {
    fprintf(stderr, "LDPC handler: error %d for epoch %08x\n", errorCode, pb->startEpoch);
    return errorCode;
  }
  errorCode = ldpc_sendProducedPackets(ctx);
  if (errorCode) return errorCode;
  if (finished) {
    /* privAmp_doPrivAmp credits one redundant bit per corrected error (priv_amp.c:111,166): that is Cascade's parity
       bookkeeping and does not hold for LDPC parities, so the credit is cancelled before the hand-over */
    pb->leakageBits += pb->correctedErrors;
    return privAmp_sendPrivAmpMsgAndPrivAmp(pb);
  }
  return 0;
}

int ldpc_onParity(ProcessBlock *pb, char *receivebuf) { return ldpc_handle(pb, receivebuf, PROC_ROLE_EC_FOLLOWER); }
int ldpc_onNack(ProcessBlock *pb, char *receivebuf) { return ldpc_handle(pb, receivebuf, PROC_ROLE_EC_INITIATOR); }
int ldpc_onMore(ProcessBlock *pb, char *receivebuf) { return ldpc_handle(pb, receivebuf, PROC_ROLE_EC_FOLLOWER); }
int ldpc_onDone(ProcessBlock *pb, char *receivebuf) { return ldpc_handle(pb, receivebuf, PROC_ROLE_EC_INITIATOR); }

int ldpc_receivePrivAmpMsg(ProcessBlock *pb, char *receivebuf) {
  /* privAmp_doPrivAmp sizes the final key from the LOCAL pb->leakageBits and ignores the `lostbits` argument it is handed
     (priv_amp.c:91,166).  Both sides must cut the same number of bits: take the initiator's figure, which already carries
     the cancellation of the Cascade-only redundancy credit (see ldpc_handle). */
  pb->leakageBits = ((EcPktHdr_StartPrivAmp *)receivebuf)->lostbits;
  return privAmp_receivePrivAmpMsg(pb, receivebuf);
}

void ldpc_releaseBlock(ProcessBlock *pb) {
  if (ldpcContext) qldpc_ecd2_release(ldpcContext, pb->startEpoch);
}
