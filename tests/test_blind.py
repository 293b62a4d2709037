"""BASELINE config 5: the ecd2 LDPC plug-in (qcrypto-ldpc_b200/host/qldpc_blind.hpp) -- blind reconciliation with
retransmission rounds over ecd2-style packets, two parties in loop-back (host/driver_blind.cpp).
GPU test: the run must agree with a model of the same protocol that decodes with the CPU oracle
(same rounds, leakage, corrected-error count, corrected key bits)."""
import json
import math
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "qcrypto-ldpc_b200", "host")
EPOCH0 = 0xb0b80000


@pytest.fixture(scope="module")
def driver(tmp_path_factory, q):
    exe = str(tmp_path_factory.mktemp("blind") / "driver_blind")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", HOST,
                           os.path.join(HOST, "driver_blind.cpp"), "-o", exe, q.LIB_PATH,
                           "-Wl,-rpath," + os.path.dirname(q.LIB_PATH)])
    return exe


def _write_keys(path, n_blocks, workbits, qber, seed):
    rng = np.random.default_rng(seed)
    words = (workbits + 31) // 32
    A = rng.integers(0, 2, (n_blocks, words * 32)).astype(np.uint8)
    A[:, workbits:] = 0
    e = (rng.random((n_blocks, words * 32)) < qber).astype(np.uint8)
    e[:, workbits:] = 0
    B = A ^ e
    from importlib import import_module
    q = import_module("qcrypto-ldpc_b200")
    with open(path, "wb") as f:
        f.write(struct.pack("<iif", n_blocks, workbits, qber))
        for b in range(n_blocks):
            f.write(q.pack_bits(A[b][None, :])[0].astype("<u4").tobytes())
            f.write(q.pack_bits(B[b][None, :])[0].astype("<u4").tobytes())
    return A, B


def h2(p):
    return -p * math.log2(p) - (1 - p) * math.log2(1 - p)


def test_blind_driver_fails_loudly_without_gpu(driver, data_dir, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    _write_keys(tmp_path / "k.bin", 1, 20000, 0.03, 1)
    p = subprocess.run([driver, "%s/NR_1_1_384.qc" % data_dir, str(tmp_path / "k.bin"), str(tmp_path / "c.bin")],
                       capture_output=True, text=True)
    assert p.returncode == 3 and "no sm_100 CUDA device" in p.stderr
    assert subprocess.run([driver], capture_output=True).returncode == 2


@pytest.mark.gpu
@pytest.mark.parametrize("qber,f_start,delta", [(0.03, 1.25, 2), (0.05, 1.05, 1), (0.08, 1.0, 3)])
def test_blind_reconciliation_matches_oracle_model(driver, q, O, data_dir, tmp_path, qber, f_start, delta):
    n_blocks, workbits, max_iter = 3, 40000, 20
    A, B = _write_keys(tmp_path / "k.bin", n_blocks, workbits, qber, seed=int(qber * 1000))
    p = subprocess.run([driver, "%s/NR_1_1_384.qc" % data_dir, str(tmp_path / "k.bin"), str(tmp_path / "c.bin"),
                        str(f_start), str(delta), str(max_iter)], capture_output=True, text=True)
    assert p.returncode == 0, (p.returncode, p.stderr, p.stdout)
    res = json.loads(p.stdout)
    words = (workbits + 31) // 32
    got = np.frombuffer(open(tmp_path / "c.bin", "rb").read(), dtype="<u4").reshape(n_blocks, words)
    got_bits = q.unpack_bits(got, words * 32)
    assert (got_bits[:, :workbits] == A[:, :workbits]).all() and res["blocks_differ"] == 0 and res["done"] == n_blocks

    # ---- the same protocol with the CPU oracle as decoder
    full = O.Code.from_qc("%s/NR_1_1_384.qc" % data_dir)
    base, Z = np.asarray(full.base), 384
    R, C = base.shape
    kcols, K = C - R, (C - R) * Z
    frames = -(-workbits // K)
    m0 = min(R, max(4, math.ceil(np.float32(f_start) * kcols * h2(float(np.float32(qber))))))
    assert res["initial_rows"] == m0 and res["frames_per_block"] == frames
    mag = min(30.0, 4.0 * math.log((1 - qber) / qber))
    mag = int(np.round(np.float32(mag)))
    codes = {}
    leak = corrected = 0
    rounds = []
    revealed = set()
    for b in range(n_blocks):
        a_pad = np.zeros(frames * K, np.uint8); a_pad[:workbits] = A[b, :workbits]
        b_pad = np.zeros(frames * K, np.uint8); b_pad[:workbits] = B[b, :workbits]
        rows = {f: m0 for f in range(frames)}
        leak += frames * m0 * Z
        todo, rnd = list(range(frames)), 0
        while todo:
            nxt = []
            for f in todo:
                m = rows[f]
                if m not in codes:
                    codes[m] = O.Code.from_base(base[:m, :kcols + m].copy(), Z)
                cw = full.nr_encode(a_pad[f * K:(f + 1) * K])
                llr = np.empty(kcols * Z + m * Z, np.int8)
                llr[:K] = np.where(b_pad[f * K:(f + 1) * K], -mag, mag)
                llr[K:] = np.where(cw[K:K + m * Z], -31, 31)
                hard, it, ok, _ = codes[m].batch_layered_fixed_i8(llr[None, :], None, rule=O.RULE_NMS, n_ite=max_iter,
                                                                  early_stop=True, norm_eighths=6)
                if ok[0]:
                    corrected += int((hard[0, :K] != b_pad[f * K:(f + 1) * K])[:max(0, min(K, workbits - f * K))].sum())
                    b_pad[f * K:(f + 1) * K] = hard[0, :K]
                else:
                    nxt.append(f)
            if nxt:
                rnd += 1
                again = []
                for f in nxt:
                    if rows[f] >= R:      # every row sent and still failing: Alice reveals the frame
                        revealed.add((b, f))
                        leak += K
                        corrected += int((a_pad[f * K:(f + 1) * K] != b_pad[f * K:(f + 1) * K]).sum())
                        b_pad[f * K:(f + 1) * K] = a_pad[f * K:(f + 1) * K]
                    else:
                        to = min(R, rows[f] + delta)
                        leak += (to - rows[f]) * Z
                        rows[f] = to
                        again.append(f)
                nxt = again
            todo = nxt
        rounds.append(rnd)
    # confirmation: one CRC-32 per frame in the clear = 32 linear parities, counted once for every frame that was not revealed
    leak += 32 * (n_blocks * frames - len(revealed))
    assert res["leak_bits"] == leak == res["leak_bits_bob"]
    assert res["corrected_errors"] == corrected == int((A ^ B)[:, :workbits].sum())
    hist = {}
    for r in rounds:
        hist[str(r)] = hist.get(str(r), 0) + 1
    assert res["round_hist"] == hist
    assert abs(res["efficiency"] - leak / (n_blocks * workbits * h2(qber))) < 1e-3


@pytest.mark.gpu
def test_blind_many_blocks_throughput_line(driver, data_dir, tmp_path):
    """a batch of 64 blocks of 65 535 bits (ecd2's block cap, processblock_mgmt.c:94): all reconciled, sane efficiency"""
    _write_keys(tmp_path / "k.bin", 64, 65535, 0.03, seed=5)
    p = subprocess.run([driver, "%s/NR_1_1_384.qc" % data_dir, str(tmp_path / "k.bin"), str(tmp_path / "c.bin")],
                       capture_output=True, text=True)
    assert p.returncode == 0, (p.stderr, p.stdout)
    res = json.loads(p.stdout)
    assert res["done"] == 64 and res["blocks_differ"] == 0
    assert 1.0 < res["efficiency"] < 2.5
    assert res["reconciled_key_bits_per_s"] > 1e6


@pytest.mark.gpu
def test_blind_crc_confirmation_catches_a_wrong_frame(driver, q, data_dir, tmp_path):
    """the confirmation step (one CRC-32 per frame in LDPC_DONE): a bit flipped in Bob's block after decoding is caught by
    Alice, the frame is revealed, and the keys end up identical; the extra leakage is the K bits of that frame minus the 32
    bits its CRC had already cost"""
    n_blocks, workbits, qber = 2, 40000, 0.03
    A, B = _write_keys(tmp_path / "k.bin", n_blocks, workbits, qber, seed=77)
    args = [driver, "%s/NR_1_1_384.qc" % data_dir, str(tmp_path / "k.bin"), str(tmp_path / "c.bin"), "1.9", "2", "20"]
    clean = json.loads(subprocess.run(args, capture_output=True, text=True, check=True).stdout)
    assert clean["crc_mismatches"] == 0
    p = subprocess.run(args, capture_output=True, text=True, env=dict(os.environ, QLDPC_BLIND_CORRUPT="1"))
    assert p.returncode == 0, (p.stderr, p.stdout)
    res = json.loads(p.stdout)
    assert res["crc_mismatches"] == 1 and res["blocks_differ"] == 0 and res["done"] == n_blocks
    assert res["leak_bits"] == clean["leak_bits"] + 22 * 384 - 32 == res["leak_bits_bob"]   # its CRC had been counted already
    words = (workbits + 31) // 32
    got = q.unpack_bits(np.frombuffer(open(tmp_path / "c.bin", "rb").read(), dtype="<u4").reshape(n_blocks, words), words * 32)
    assert (got[:, :workbits] == A[:, :workbits]).all()


@pytest.fixture(scope="module")
def packet_tester(tmp_path_factory, q):
    exe = str(tmp_path_factory.mktemp("blindpk") / "test_blind_packets")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", HOST,
                           os.path.join(HOST, "test_blind_packets.cpp"), "-o", exe, q.LIB_PATH,
                           "-Wl,-rpath," + os.path.dirname(q.LIB_PATH)])
    return exe


def test_bob_rejects_malformed_packets_without_touching_memory(packet_tester, data_dir):
    """every field of a received packet is checked against totalLengthInBytes and the block before anything is
    copied (lengths, z, rows <= R, frame indices, row ranges); runs on the CPU -- rejection happens before any decode"""
    p = subprocess.run([packet_tester, "%s/NR_1_1_384.qc" % data_dir, "bob"], capture_output=True, text=True)
    assert p.returncode == 0, (p.stdout, p.stderr)
    assert "bob: all malformed packets rejected" in p.stdout


@pytest.mark.gpu
def test_alice_rejects_malformed_packets(packet_tester, data_dir):
    p = subprocess.run([packet_tester, "%s/NR_1_1_384.qc" % data_dir, "alice"], capture_output=True, text=True)
    assert p.returncode == 0, (p.stdout, p.stderr)
    assert "alice: all malformed packets rejected" in p.stdout
