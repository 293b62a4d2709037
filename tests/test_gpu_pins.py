"""The CUDA path against the results the reference RECORDED for its own decoders, and the 10^4-frame bit-exact gate of
SURVEY.md 7.5.  Everything goes through the C ABI; the oracle is the checker on a subset (bit-exact), the recorded tables
(tests/golden/recorded_fer.json, workloads in tests/refpins.py) are the statistical pin."""
import numpy as np
import pytest

import refpins
from conftest import make_frames

pytestmark = pytest.mark.gpu


def _gpu_nr_encoder(q, code):
    """qldpc_encode_nr as the test's Alice (itself checked against the oracle in test_gpu_flooding / test_abi)"""
    enc = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=q.RULE_NMS, dtype=q.DTYPE_F32, max_iter=1)
    return enc, (lambda msgs: q.unpack_bits(enc.encode_nr(q.pack_bits(msgs)), code.n))


@pytest.mark.parametrize("name", ["NR_1_1_24", "NR_2_6_52"])
def test_matlab_rm_fp_recorded_table_on_gpu(q, O, data_dir, name):
    """BPSK_nrldpc_sim_RM_FP.m:1-37 exactly (rmax 3, maxqr 31, maxqL 127, offset 2, 20 iterations,
    no early stop, nbRM = ceil(kb / Rate) + 2, first 2z LLRs zeroed, floor quantiser), 20 000 frames per Eb/N0 point on the
    GPU, the oracle bit-exact on a 1 500-frame subset of every point; FER against ML/sim_results.m (k/n rate definition,
    see refpins.py): every point statistically consistent, at least 3 of the 4 estimates inside the recorded 95 % interval."""
    rows = refpins.recorded()["sim_results_m"][name]
    full = q.Code.from_qc_file("%s/%s.qc" % (data_dir, name))
    ofull = O.Code.from_qc("%s/%s.qc" % (data_dir, name))
    kb, nbRM, mbRM = refpins.rm_fp_geometry(ofull.base)
    sub_base = ofull.base[:mbRM, :nbRM].copy()
    sub = q.Code.from_qc(sub_base, ofull.Z)
    osub = O.Code.from_base(sub_base, ofull.Z)
    enc, encode = _gpu_nr_encoder(q, full)
    dec = q.Decoder(sub, schedule=q.SCHED_LAYERED, rule=q.RULE_OMS, dtype=q.DTYPE_I8, max_iter=20, early_stop=False,
                    offset=2.0, msg_max=31, app_max=127, out_mode=q.OUT_INFO)
    assert dec.kernel_name == "layered_i8_zpack4"
    F, F_ORACLE = 20000, 1500
    inside = 0
    for row in rows:
        msgs, llr = refpins.rm_fp_frames(ofull.base, ofull.Z, row["ebno_db"], F, encode, seed=1)
        out, ok, iters, _ = dec.decode(llr)
        got = q.unpack_bits(out, sub.k)
        hard, oit, ook, _ = osub.batch_layered_fixed_i8(llr[:F_ORACLE], None, rule=O.RULE_OMS, n_ite=20, early_stop=False,
                                                        offset=2, msg_max=31, app_max=127)
        assert (got[:F_ORACLE] == hard[:, :sub.k]).all() and (iters[:F_ORACLE] == oit).all() and (ok[:F_ORACLE] == ook).all()
        fe = int((got != msgs).any(axis=1).sum())
        assert refpins.consistent(fe, F, row["frame_errors"], row["frames"]), (row, fe)
        lo, hi = refpins.clopper_pearson(row["frame_errors"], row["frames"])
        inside += lo <= fe / F <= hi
        print("%s %.2f dB: gpu FER %.5f (%d / %d), recorded %.5f [%.5f, %.5f]" % (name, row["ebno_db"], fe / F, fe, F,
                                                                                   row["fer"], lo, hi))
    assert inside >= 3
    dec.close()
    enc.close()


def test_5g_qc_driver_recorded_table_on_gpu(q, O, data_dir):
    """flooding SPA on NR_1_0_2.qc with the (5g-qc) puncture rule ("main.cpp (5g-qc)":450-457,
    514-530), 4 000 GPU frames per QBER step against README_LDPC.md:941-974 (100 frames each): every step statistically
    consistent, at least 9 of 11 estimates inside the recorded 95 % interval, BER of the sweep within 15 %; decoded bits
    equal to the oracle's and posteriors within 1e-3 on a 1 000-frame subset of every step."""
    rows = [r for r in refpins.recorded()["readme_ldpc_nr_1_0_2"]["rows"] if r["qber"] > 0]
    path = "%s/NR_1_0_2.qc" % data_dir
    code = q.Code.from_qc_file(path)
    oc = O.Code.from_qc(path)
    enc, encode = _gpu_nr_encoder(q, code)
    dec = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=q.RULE_SPA, dtype=q.DTYPE_F32, max_iter=10, early_stop=True,
                    syndrome_depth=1, out_mode=q.OUT_ALL)
    F, F_ORACLE = 4000, 1000
    inside = 0
    be_sum = be_ref = 0.0
    for row in rows:
        msgs, llr, punct = refpins.qc5g_frames(oc.N, oc.K, row["qber"], F, encode, seed=int(row["qber"] * 100))
        assert punct == row["punctured_bits"]
        out, ok, iters, post = dec.decode(llr, want_posterior=True)
        got = q.unpack_bits(out, oc.N)
        hard, opost, oit, ook, _ = oc.batch_flooding_f32(llr[:F_ORACLE], None, rule=O.RULE_SPA, n_ite=10, early_stop=True)
        assert (got[:F_ORACLE] == hard).all() and (iters[:F_ORACLE] == oit).all() and (ok[:F_ORACLE] == ook).all()
        np.testing.assert_allclose(post[:F_ORACLE], opost, rtol=1e-3, atol=1e-3)
        fe = int((got[:, :oc.K] != msgs).any(axis=1).sum())
        assert refpins.consistent(fe, F, row["frame_errors"], row["frames"]), (row, fe)
        lo, hi = refpins.clopper_pearson(row["frame_errors"], row["frames"])
        inside += lo <= fe / F <= hi
        be_sum += float((got[:, :oc.K] != msgs).mean())
        be_ref += row["ber"]
        print("QBER %.2f: gpu FER %.4f, recorded %.2f [%.3f, %.3f]" % (row["qber"], fe / F, row["fer"], lo, hi))
    assert inside >= 9
    assert abs(be_sum / be_ref - 1) < 0.15
    dec.close()
    enc.close()


@pytest.mark.parametrize("mode,rule,early", [("parity", "nms", True), ("syndrome", "oms", True), ("parity", "nms", False)])
def test_bg1_z384_16384_frames_bit_exact(q, O, data_dir, mode, rule, early):
    """VERDICT r1 item 1(c) / SURVEY 7.5: a 16 384-frame BG1 Z=384 batch, EVERY frame's bits, syndrome flag and iteration
    count equal to the oracle's -- send-parity and syndrome formulation, early stop and fixed iterations."""
    path = "%s/NR_1_1_384.qc" % data_dir
    oc = O.Code.from_qc(path)
    F = 16384
    rng = np.random.default_rng(20261018)
    qber = 0.03 if mode == "parity" else 0.045
    if mode == "parity":
        code = q.Code.from_qc_file(path)
        enc, encode = _gpu_nr_encoder(q, code)
        msgs = rng.integers(0, 2, (F, oc.K)).astype(np.uint8)
        cw = encode(msgs)
        enc.close()
        llr = np.where(cw, -31, 31).astype(np.int8)
        e = rng.random((F, oc.K)) < qber
        llr[:, :oc.K] = np.where(cw[:, :oc.K] ^ e, -14, 14)
        syn = None
    else:
        x = rng.integers(0, 2, (F, oc.N)).astype(np.uint8)
        code = q.Code.from_qc_file(path)
        tmp = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=1, norm_factor=0.75)
        syn_packed = tmp.syndrome(q.pack_bits(x))
        tmp.close()
        syn = q.unpack_bits(syn_packed, oc.M)
        llr = np.where(x ^ (rng.random((F, oc.N)) < qber), -12, 12).astype(np.int8)
    r = q.RULE_NMS if rule == "nms" else q.RULE_OMS
    n_ite = 10 if early else 4
    dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=r, dtype=q.DTYPE_I8, max_iter=n_ite, early_stop=early,
                    norm_factor=0.75, offset=2.0, out_mode=q.OUT_ALL)
    assert dec.kernel_name == "layered_i8s_zpack4"
    out, ok, iters, _ = dec.decode(llr, None if syn is None else q.pack_bits(syn))
    got = q.unpack_bits(out, oc.N)
    hard, oit, ook, _ = oc.batch_layered_fixed_i8(llr, syn, rule=O.RULE_NMS if rule == "nms" else O.RULE_OMS, n_ite=n_ite,
                                                  early_stop=early, offset=2, norm_eighths=6)
    assert (iters == oit).all() and (ok == ook).all()
    assert (got == hard).all()
    if early:
        assert ok.mean() > 0.99 and len(set(iters.tolist())) >= 2      # the batch exercises several iteration counts
    dec.close()
