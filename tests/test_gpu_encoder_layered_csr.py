"""SURVEY.md 8 rows a6 and a9 for codes that are NOT quasi-cyclic (the .alist matrices the reference's drivers load):
the systematic encoders (Encoder_LDPC_from_H, Encoder_LDPC with a G file) against the reference's known-answer vector
KAT-E, and the horizontal-layered schedule on an arbitrary H against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_kat_e_through_both_encoders(q, O, data_dir, kat):
    """data[504] -> encoded[1008] ("main.cpp (alist)":443-455): from H alone (Gauss-Jordan keeps the information positions
    504..1007 of BOOT/matrices/G/PEGReg504x1008.alist) and from the reference's generator-matrix file"""
    code = q.Code.from_alist("%s/PEGReg504x1008.alist" % data_dir)
    data = np.array(kat["data"], dtype=np.uint8)
    want = np.array(kat["encoded"], dtype=np.uint8)
    for enc in (q.Encoder.from_H(code), q.Encoder.from_G_alist("%s/G_PEGReg504x1008.alist" % data_dir)):
        assert (enc.k, enc.n) == (504, 1008)
        assert enc.info_bits_pos.tolist() == list(range(504, 1008))
        rng = np.random.default_rng(2)
        msgs = np.concatenate([data[None, :], rng.integers(0, 2, (300, 504)).astype(np.uint8)])
        cw = q.unpack_bits(enc.encode(q.pack_bits(msgs)), 1008)
        assert (cw[0] == want).all()
        assert (cw[:, 504:] == msgs).all()
        oc = O.Code.from_alist("%s/PEGReg504x1008.alist" % data_dir)
        assert not any(oc.syndrome(cw[f]).any() for f in range(0, 301, 25))
        enc.close()


@pytest.mark.parametrize("name", ["20.alist", "1998.5.3.2665.alist"])
def test_encoder_from_h_on_the_other_alist_codes(q, O, data_dir, name):
    """irregular dv 2..20 (20.alist) and the (4,36)-regular rate-8/9 code: whatever information positions the elimination
    ends with, every codeword satisfies H, is systematic in them, and the map is linear"""
    code = q.Code.from_alist("%s/%s" % (data_dir, name))
    oc = O.Code.from_alist("%s/%s" % (data_dir, name))
    enc = q.Encoder.from_H(code)
    assert enc.n == oc.N and enc.k >= oc.N - oc.M
    rng = np.random.default_rng(4)
    msgs = rng.integers(0, 2, (64, enc.k)).astype(np.uint8)
    cw = q.unpack_bits(enc.encode(q.pack_bits(msgs)), enc.n)
    assert (cw[:, enc.info_bits_pos] == msgs).all()
    assert not any(oc.syndrome(cw[f]).any() for f in range(64))
    s = q.unpack_bits(enc.encode(q.pack_bits(msgs[:32] ^ msgs[32:])), enc.n)
    assert (s == (cw[:32] ^ cw[32:])).all()
    enc.close()


@pytest.mark.parametrize("name,qber", [("PEGReg504x1008.alist", 0.04), ("20.alist", 0.03)])
@pytest.mark.parametrize("rule", ["spa", "nms", "oms"])
def test_layered_schedule_on_an_arbitrary_h(q, O, data_dir, name, qber, rule):
    """Decoder_LDPC_BP_horizontal_layered takes any Sparse_matrix ("main.cpp (5g-qc)":256-270): row-serial layered BP on .alist
    codes, one thread per frame; bits, iteration counts, flags equal to the oracle's, posteriors within 1e-3"""
    path = "%s/%s" % (data_dir, name)
    oc = O.Code.from_alist(path)
    code = q.Code.from_alist(path)
    qr, orr = {"spa": (q.RULE_SPA, O.RULE_SPA), "nms": (q.RULE_NMS, O.RULE_NMS), "oms": (q.RULE_OMS, O.RULE_OMS)}[rule]
    dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=qr, dtype=q.DTYPE_F32, max_iter=15, early_stop=True, norm_factor=0.8125,
                    offset=0.25, out_mode=q.OUT_ALL)
    assert dec.kernel_name == "layered_csr"
    F = 300
    rng = np.random.default_rng(6)
    x = rng.integers(0, 2, (F, oc.N)).astype(np.uint8)
    y = x ^ (rng.random((F, oc.N)) < qber)
    syn = np.stack([oc.syndrome(x[f]) for f in range(F)])
    mag = np.float32(np.log((1 - qber) / qber))
    llr = np.where(y, -mag, mag).astype(np.float32)
    out, ok, iters, post = dec.decode(llr, q.pack_bits(syn), want_posterior=True)
    got = q.unpack_bits(out, oc.N)
    for f in range(F):
        h, p, it, o = oc.decode_layered_f32(llr[f], syn[f], rule=orr, n_ite=15, early_stop=True, norm=0.8125, offset=0.25)
        assert (got[f] == h).all() and iters[f] == it and ok[f] == o, f
        np.testing.assert_allclose(post[f], p, rtol=1e-3, atol=1e-4)
    assert ok.mean() > 0.5 and len(set(iters.tolist())) >= 3
    st = dec.stats()
    assert st["frames"] == F and st["iter_sum"] == int(iters.sum())
    # integer dtypes have no definition on a non-quasi-cyclic H
    with pytest.raises(q.QldpcError) as e:
        q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=5, norm_factor=0.75)
    assert e.value.code == 6
    dec.close()
