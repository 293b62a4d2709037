/**
 * @file ldpc_reconcile.h
 * @brief Blind LDPC reconciliation for ecd2 (algorithm slots ALG_LDPC_CONTINUE_ROLES / ALG_LDPC_FLIP_ROLES), decoded on a
 *        B200 through libqldpc_b200.so (include/qldpc_ecd2.h of the qcrypto-ldpc_b200 repository).
 *
 * Added by integration/ecd2_ldpc.patch.  Packet subtypes 9..12 (definitions/packets.h), handler arrays
 * ALG_PKTHNDLRS_LDPC_INITIATOR / _FOLLOWER and ALG_DATA_MNGR_LDPC (definitions/algorithms/algorithms.c).
 *
 * Configuration comes from the environment, read once when the first block reaches error correction:
 *   ECD2_LDPC_BASE_QC   NR base graph file (.qc), mandatory
 *   ECD2_LDPC_DEVICE    CUDA ordinal (default 0)
 *   ECD2_LDPC_F_START, ECD2_LDPC_DELTA_ROWS, ECD2_LDPC_MAX_ITER   protocol parameters (defaults 1.25, 2, 20)
 * Without an sm_100 device (or without the base graph) every entry point returns 81, as the reference does today.
 */
#ifndef ECD2_LDPC_RECONCILE
#define ECD2_LDPC_RECONCILE

#include "../definitions/processblock.h"

/// Prepares a block for LDPC error correction on either side: removes the bits revealed by the QBER estimation
/// (helper_cleanupRevealedBits: sets workbits, clears leakageBits) -- what helper_prepPermutationWrapper does for Cascade.
void ldpc_prepareBlock(ProcessBlock *pb);
/// EC initiator: encode the block, send the first parity rows (subtype 9).  Counterpart of cascade_initiateAfterQber.
int ldpc_initiateAfterQber(ProcessBlock *pb);
/// Packet handlers (PacketHandlerArray entries): 9 parity rows / 11 more rows or revealed frames -> follower;
/// 10 failed-frame list / 12 per-frame CRCs -> initiator.  A handler that receives a subtype meant for the other role returns 45.
int ldpc_onParity(ProcessBlock *pb, char *receivebuf);
int ldpc_onNack(ProcessBlock *pb, char *receivebuf);
int ldpc_onMore(ProcessBlock *pb, char *receivebuf);
int ldpc_onDone(ProcessBlock *pb, char *receivebuf);
/// Subtype 8 on an LDPC block: privAmp_receivePrivAmpMsg with the initiator's leakage figure installed first
int ldpc_receivePrivAmpMsg(ProcessBlock *pb, char *receivebuf);
/// drops the per-block protocol state kept in the library (called by freeLdpcData)
void ldpc_releaseBlock(ProcessBlock *pb);

#endif
