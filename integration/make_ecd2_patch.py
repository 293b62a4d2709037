#!/usr/bin/env python3
"""Regenerates integration/ecd2_ldpc.patch from the read-only reference checkout (build container only).

    python integration/make_ecd2_patch.py

The patch registers the blind LDPC reconciliation in the reference's ecd2 daemon (errorcorrection/):
  definitions/packets.h                  EC subtypes 9..12, EC_SUBTYPE_COUNT                       (:46-56)
  subcomponents/comms.c                  sizes of the new headers in comms_createEcHeader            (:153-168)
  definitions/algorithms/data_manager.h  ALG_DATATYPE_LDPC, LdpcData                                 (:17-20)
  definitions/algorithms/algorithms.h/.c handler arrays, packet managers, data manager, initDataStruct case (:62-94,121-132)
  subcomponents/qber_estim.c             the two `return 81` bodies of ALG_LDPC_*; algorithm choice by environment (:301,337-340,420-423)
  ecd2.h                                 errormessage[85]                                             (:251-337)
  ecd2.c                                 the dispatch loop keeps the handler's return value          (:525-526)
  Makefile                               ldpc_reconcile.o, -lqldpc_b200
  subcomponents/ldpc_reconcile.{h,c}     new files (integration/ecd2/)
Each edit is a search-and-replace on the reference's own text, so the script fails loudly if the reference changes.
"""
import os
import shutil
import subprocess
import tempfile

REF = "/root/reference/errorcorrection"
HERE = os.path.dirname(os.path.abspath(__file__))


def edit(path, pairs, append=""):
    """search-and-replace that keeps the file's own line endings (several reference files are CRLF)"""
    s = open(path, newline="").read()
    crlf = "\r\n" in s
    for old, new in pairs:
        if crlf:
            old, new = old.replace("\n", "\r\n"), new.replace("\n", "\r\n")
        assert s.count(old) == 1, (path, old[:60], s.count(old))
        s = s.replace(old, new)
    s += append.replace("\n", "\r\n") if crlf else append
    open(path, "w", newline="").write(s)


def main():
    tmp = tempfile.mkdtemp(prefix="ecd2patch")
    a, b = os.path.join(tmp, "a", "errorcorrection"), os.path.join(tmp, "b", "errorcorrection")
    ignore = shutil.ignore_patterns("ldpc_examples", "readme_imgs", "*.md", "LICENSE")
    shutil.copytree(REF, a, ignore=ignore)
    shutil.copytree(REF, b, ignore=ignore)
    for f in ("ldpc_reconcile.h", "ldpc_reconcile.c"):
        shutil.copyfile(os.path.join(HERE, "ecd2", f), os.path.join(b, "subcomponents", f))

    edit(b + "/definitions/packets.h", [
        ("    SUBTYPE_START_PRIV_AMP = 8,\n};\n#define EC_SUBTYPE_COUNT 8 ///< Keep this updated \n",
         "    SUBTYPE_START_PRIV_AMP = 8,\n"
         "    SUBTYPE_LDPC_PARITY = 9,   ///< EC initiator -> follower: first parity block rows of every frame\n"
         "    SUBTYPE_LDPC_NACK = 10,    ///< follower -> initiator: frames whose syndrome check failed\n"
         "    SUBTYPE_LDPC_MORE = 11,    ///< initiator -> follower: more parity rows, or the revealed key bits of a frame\n"
         "    SUBTYPE_LDPC_DONE = 12,    ///< follower -> initiator: one CRC-32 per corrected frame\n"
         "};\n#define EC_SUBTYPE_COUNT 12 ///< Keep this updated \n"),
        ("/**\n * @brief structure to hold received messages\n",
         "/// @name LDPC reconciliation (subtypes 9..12); layouts shared with libqldpc_b200 (host/qldpc_blind.hpp)\n/// @{\n"
         "typedef struct ERRC_LDPC_9 {\n    EcPktHdr_Base base;\n"
         "    unsigned int z, frames, first_frame, rows; /**< lifting size; frames in this packet; parity block rows 0..rows-1 follow */\n"
         "    unsigned int workbits;\n    float qber;\n} EcPktHdr_LdpcParity;\n"
         "typedef struct ERRC_LDPC_10 {\n    EcPktHdr_Base base;\n    unsigned int round, n_failed; /**< n_failed frame indices follow */\n} EcPktHdr_LdpcNack;\n"
         "typedef struct ERRC_LDPC_11 {\n    EcPktHdr_Base base;\n    unsigned int round, n_frames, row_from, row_to, reveal;\n} EcPktHdr_LdpcMore;\n"
         "typedef struct ERRC_LDPC_12 {\n    EcPktHdr_Base base;\n    unsigned int rounds, frames_revealed, frames, corrected_errors; /**< `frames` CRC-32 values follow */\n} EcPktHdr_LdpcDone;\n"
         "/// @}\n\n/**\n * @brief structure to hold received messages\n"),
    ])
    edit(b + "/subcomponents/comms.c", [
        ("    case SUBTYPE_START_PRIV_AMP:              size += sizeof(EcPktHdr_StartPrivAmp);            break;\n",
         "    case SUBTYPE_START_PRIV_AMP:              size += sizeof(EcPktHdr_StartPrivAmp);            break;\n"
         "    case SUBTYPE_LDPC_PARITY:                 size += sizeof(EcPktHdr_LdpcParity);              break;\n"
         "    case SUBTYPE_LDPC_NACK:                   size += sizeof(EcPktHdr_LdpcNack);                break;\n"
         "    case SUBTYPE_LDPC_MORE:                   size += sizeof(EcPktHdr_LdpcMore);                break;\n"
         "    case SUBTYPE_LDPC_DONE:                   size += sizeof(EcPktHdr_LdpcDone);                break;\n"),
    ])
    edit(b + "/definitions/algorithms/data_manager.h", [
        ("    ALG_DATATYPE_CASCADE,\n} ALGORITHM_DATATYPE;", "    ALG_DATATYPE_CASCADE,\n    ALG_DATATYPE_LDPC,\n} ALGORITHM_DATATYPE;"),
        ("} CascadeData;\n", "} CascadeData;\n\n/**\n * @brief struct for the LDPC reconciliation; the per-frame protocol state lives in libqldpc_b200, keyed by epoch\n * \n */\n"
                             "typedef struct ALGORITHM_LDPC_DATA {\n    int blockRegistered;                  /**< Boolean. the library holds state for this block */\n} LdpcData;\n"),
    ])
    edit(b + "/definitions/algorithms/algorithms.h", [
        ('#include "../../subcomponents/qber_estim.h"\n', '#include "../../subcomponents/qber_estim.h"\n#include "../../subcomponents/ldpc_reconcile.h"\n'),
        ("extern const PacketHandlerArray ALG_PKTHNDLRS_CASCADE_INITIATOR;\n",
         "extern const PacketHandlerArray ALG_PKTHNDLRS_CASCADE_INITIATOR;\nextern const PacketHandlerArray ALG_PKTHNDLRS_LDPC_FOLLOWER;\nextern const PacketHandlerArray ALG_PKTHNDLRS_LDPC_INITIATOR;\n"),
        ("extern const ALGORITHM_PKT_MNGR ALG_PKT_MNGR_CASCADE_INITIATOR;\n",
         "extern const ALGORITHM_PKT_MNGR ALG_PKT_MNGR_CASCADE_INITIATOR;\nextern const ALGORITHM_PKT_MNGR ALG_PKT_MNGR_LDPC_FOLLOWER;\nextern const ALGORITHM_PKT_MNGR ALG_PKT_MNGR_LDPC_INITIATOR;\n"),
        ("extern const ALGORITHM_DATA_MNGR ALG_DATA_MNGR_CASCADE;\n", "extern const ALGORITHM_DATA_MNGR ALG_DATA_MNGR_CASCADE;\nextern const ALGORITHM_DATA_MNGR ALG_DATA_MNGR_LDPC;\n"),
    ])
    edit(b + "/definitions/algorithms/algorithms.c", [
        ("// ALGORITHM DATA STRUCT FOR PROCESS BLOCK\n",
         "/// @name LDPC PACKET MANAGERS (subtypes 8..12; 8 = privacy amplification, as for cascade)\n/// @{\n"
         "const PacketHandlerArray ALG_PKTHNDLRS_LDPC_INITIATOR = {\n"
         "    ldpc_receivePrivAmpMsg,                     ///< Subtype 8\n"
         "    ldpc_onParity,                              ///< Subtype 9  (follower only: answers 45 here)\n"
         "    ldpc_onNack,                                ///< Subtype 10\n"
         "    ldpc_onMore,                                ///< Subtype 11 (follower only)\n"
         "    ldpc_onDone                                 ///< Subtype 12\n};\n"
         "const ALGORITHM_PKT_MNGR ALG_PKT_MNGR_LDPC_INITIATOR = {\n    &ALG_PKTHNDLRS_LDPC_INITIATOR,              // funcHandlers\n"
         "    8,                                          // FIRST_SUBTYPE\n"
         "    8 + sizeof(ALG_PKTHNDLRS_LDPC_INITIATOR)    // LAST_SUBTYPE, automatically calculated\n"
         "        / sizeof(ALG_PKTHNDLRS_LDPC_INITIATOR[0]) - 1,\n    False                                       // allowNullPrcBlks\n};\n"
         "const PacketHandlerArray ALG_PKTHNDLRS_LDPC_FOLLOWER = {\n"
         "    ldpc_receivePrivAmpMsg,                     ///< Subtype 8\n"
         "    ldpc_onParity,                              ///< Subtype 9\n"
         "    ldpc_onNack,                                ///< Subtype 10 (initiator only: answers 45 here)\n"
         "    ldpc_onMore,                                ///< Subtype 11\n"
         "    ldpc_onDone                                 ///< Subtype 12 (initiator only)\n};\n"
         "const ALGORITHM_PKT_MNGR ALG_PKT_MNGR_LDPC_FOLLOWER = {\n    &ALG_PKTHNDLRS_LDPC_FOLLOWER,               // funcHandlers\n"
         "    8,                                          // FIRST_SUBTYPE\n"
         "    8 + sizeof(ALG_PKTHNDLRS_LDPC_FOLLOWER)     // LAST_SUBTYPE, automatically calculated\n"
         "        / sizeof(ALG_PKTHNDLRS_LDPC_FOLLOWER[0]) - 1,\n    False                                       // allowNullPrcBlks\n};\n/// @}\n\n"
         "// ALGORITHM DATA STRUCT FOR PROCESS BLOCK\n"),
        ("            processBlock->algorithmDataPtr = malloc2(sizeof(CascadeData));\n            break;\n",
         "            processBlock->algorithmDataPtr = malloc2(sizeof(CascadeData));\n            break;\n"
         "        case ALG_DATATYPE_LDPC:\n            processBlock->algorithmDataPtr = malloc2(sizeof(LdpcData));\n            break;\n"),
    ])
    edit(b + "/definitions/algorithms/algorithms.c", [], append=("\n\n/// @name LDPC DATA MANAGER\n/// @{\n/// @brief Function to initialize LDPC specific data\n"
                "int initLdpcData(ProcessBlock* processBlock) {\n    int errorCode = initDataStruct(processBlock);\n    if (errorCode) return errorCode;\n"
                "    ((LdpcData *)(processBlock->algorithmDataPtr))->blockRegistered = 1;\n    return 0;\n}\n"
                "/// @brief function to free LDPC specific data: the library drops its per-block protocol state\n"
                "int freeLdpcData(ProcessBlock* processBlock) {\n    ldpc_releaseBlock(processBlock);\n    return freeDataStruct(processBlock);\n}\n"
                "/// @brief Contains information & functions on the LDPC algorithm w.r.t. data handling\n"
                "const ALGORITHM_DATA_MNGR ALG_DATA_MNGR_LDPC = {\n    ALG_DATATYPE_LDPC,      // Enum which identifies the data type of the data\n"
                "    &initLdpcData,          // Function to call to initialize the data\n    &freeLdpcData           // Function to call to free the data\n};\n/// @} \n"))
    follower_old = ("    case ALG_LDPC_CONTINUE_ROLES:\n      return 81;\n    case ALG_LDPC_FLIP_ROLES:\n      return 81;\n    default:\n"
                    "      fprintf(stderr, \"Err 81 at chooseEcAlgorithmAsQberFollower\\n\");")
    follower_new = ("    case ALG_LDPC_CONTINUE_ROLES:\n      processBlock->processorRole = PROC_ROLE_EC_FOLLOWER;\n"
                    "      processBlock->algorithmPktMngr = (ALGORITHM_PKT_MNGR *)&ALG_PKT_MNGR_LDPC_FOLLOWER;\n"
                    "      processBlock->algorithmDataMngr = (ALGORITHM_DATA_MNGR *)&ALG_DATA_MNGR_LDPC;\n"
                    "      errorCode = processBlock->algorithmDataMngr->initData(processBlock);\n      if (errorCode) return errorCode;\n"
                    "      ldpc_prepareBlock(processBlock);\n"
                    "      // Insert the packet, then await the parity rows (subtype 9) from the EC_INITIATOR\n"
                    "      return comms_insertSendPacket((char *)(bufferToSend), bufferLengthInBytes);\n"
                    "    case ALG_LDPC_FLIP_ROLES:\n      processBlock->processorRole = PROC_ROLE_EC_INITIATOR;\n"
                    "      processBlock->algorithmPktMngr = (ALGORITHM_PKT_MNGR *)&ALG_PKT_MNGR_LDPC_INITIATOR;\n"
                    "      processBlock->algorithmDataMngr = (ALGORITHM_DATA_MNGR *)&ALG_DATA_MNGR_LDPC;\n"
                    "      errorCode = processBlock->algorithmDataMngr->initData(processBlock);\n      if (errorCode) return errorCode;\n"
                    "      errorCode = comms_insertSendPacket((char *)(bufferToSend), bufferLengthInBytes);\n      if (errorCode)\n        return errorCode;\n"
                    "      return ldpc_initiateAfterQber(processBlock);\n    default:\n"
                    "      fprintf(stderr, \"Err 81 at chooseEcAlgorithmAsQberFollower\\n\");")
    initiator_old = ("    case ALG_LDPC_CONTINUE_ROLES:\n      return 81;\n    case ALG_LDPC_FLIP_ROLES:\n      return 81;\n    default:\n"
                     "      fprintf(stderr, \"Err 81 at qber_prepareErrorCorrection\\n\");")
    initiator_new = ("    case ALG_LDPC_CONTINUE_ROLES:\n      processBlock->processorRole = PROC_ROLE_EC_INITIATOR;\n"
                     "      processBlock->algorithmPktMngr = (ALGORITHM_PKT_MNGR *)&ALG_PKT_MNGR_LDPC_INITIATOR;\n"
                     "      processBlock->algorithmDataMngr = (ALGORITHM_DATA_MNGR *)&ALG_DATA_MNGR_LDPC;\n"
                     "      errorCode = processBlock->algorithmDataMngr->initData(processBlock);\n      if (errorCode) return errorCode;\n"
                     "      return ldpc_initiateAfterQber(processBlock);\n"
                     "    case ALG_LDPC_FLIP_ROLES:\n      // QBER_INITIATOR is now the EC_FOLLOWER: await the parity rows (subtype 9)\n"
                     "      processBlock->processorRole = PROC_ROLE_EC_FOLLOWER;\n"
                     "      processBlock->algorithmPktMngr = (ALGORITHM_PKT_MNGR *)&ALG_PKT_MNGR_LDPC_FOLLOWER;\n"
                     "      processBlock->algorithmDataMngr = (ALGORITHM_DATA_MNGR *)&ALG_DATA_MNGR_LDPC;\n"
                     "      errorCode = processBlock->algorithmDataMngr->initData(processBlock);\n      if (errorCode) return errorCode;\n"
                     "      ldpc_prepareBlock(processBlock);\n      return 0;\n    default:\n"
                     "      fprintf(stderr, \"Err 81 at qber_prepareErrorCorrection\\n\");")
    edit(b + "/subcomponents/qber_estim.c", [
        ("  chosenAlgorithm = ALG_CASCADE_CONTINUE_ROLES;\n  // chosenAlgorithm = ALG_CASCADE_FLIP_ROLES;\n",
         "  chosenAlgorithm = ALG_CASCADE_CONTINUE_ROLES;\n  // chosenAlgorithm = ALG_CASCADE_FLIP_ROLES;\n"
         "  // ECD2_EC_ALGORITHM=3 (ALG_LDPC_CONTINUE_ROLES) or 4 (ALG_LDPC_FLIP_ROLES) selects the LDPC reconciliation\n"
         "  if (getenv(\"ECD2_EC_ALGORITHM\")) chosenAlgorithm = (ALGORITHM_DECISION)atoi(getenv(\"ECD2_EC_ALGORITHM\"));\n"),
        (follower_old, follower_new),
        (initiator_old, initiator_new),
    ])
    # the dispatch loop drops the handler's error code (ecd2.c:525-526); the LDPC handlers report through it
    edit(b + "/ecd2.c", [
        ("              (*(tmpPrcBlk->algorithmPktMngr->FUNC_HANDLERS))\n", "              errorCode = (*(tmpPrcBlk->algorithmPktMngr->FUNC_HANDLERS))\n"),
    ])
    edit(b + "/ecd2.h", [
        ("    \"Algorithm specific data ptr not null\"\n};", "    \"Algorithm specific data ptr not null\",\n    \"LDPC packet inconsistent with its block\" /* 85 */\n};"),
    ])
    edit(b + "/Makefile", [
        ("all:	ecd2 \n", "all:	ecd2 \n\n# LDPC reconciliation on a B200: root of the qcrypto-ldpc_b200 checkout (include/qldpc_ecd2.h, libqldpc_b200.so)\n"
                         "QLDPC_ROOT ?= ../qcrypto-ldpc_b200-repo\nQLDPC_LIBDIR ?= $(QLDPC_ROOT)/qcrypto-ldpc_b200\n"),
        ("# ecd2\necd2.o: ecd2.c\n", "ldpc_reconcile.o: subcomponents/ldpc_reconcile.c\n	gcc -Wall -O3 -c -g -I$(QLDPC_ROOT)/include subcomponents/ldpc_reconcile.c\n\n# ecd2\necd2.o: ecd2.c\n"),
        ("		priv_amp.o qber_estim.o processblock_mgmt.o algorithms.o ecd2.o\n		\n",
         "		priv_amp.o qber_estim.o processblock_mgmt.o algorithms.o ldpc_reconcile.o ecd2.o\n		\n"),
        ("		processblock_mgmt.o algorithms.o ecd2.o -lm\n",
         "		processblock_mgmt.o algorithms.o ldpc_reconcile.o ecd2.o -lm \\\n		-L$(QLDPC_LIBDIR) -lqldpc_b200 -Wl,-rpath,$(abspath $(QLDPC_LIBDIR))\n"),
    ])
    out = os.path.join(HERE, "ecd2_ldpc.patch")
    p = subprocess.run(["diff", "-ruN", "a/errorcorrection", "b/errorcorrection"], cwd=tmp, capture_output=True)
    assert p.returncode == 1, p.stderr
    # bytes, not text: several reference files are CRLF and the hunks must keep that; timestamps out of the headers so
    # that the patch is reproducible
    lines = []
    for ln in p.stdout.split(b"\n"):
        if ln.startswith(b"--- ") or ln.startswith(b"+++ "):
            ln = ln.split(b"\t")[0]
        if ln.startswith(b"diff -ruN"):
            continue
        lines.append(ln)
    open(out, "wb").write(b"\n".join(lines))
    shutil.rmtree(tmp)
    print("wrote", out, "(%d lines)" % len(lines))


if __name__ == "__main__":
    main()
