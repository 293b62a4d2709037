"""Parity of the CUDA flooding decoders (through the C ABI) against the CPU oracle and the reference's
known-answer vector.  Float tier: decoded bits equal, posteriors within 1e-3 relative (north_star);
integer tiers: bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-3   # tolerance on posterior LLRs stated by BASELINE.json north_star


def _bsc_llr_frames(oc, F, qber, seed, confirmed=None):
    """the reference's chain: random codeword of the zero-syndrome coset is not needed for a symmetric
    decoder, so frames are x (random), syndrome H*x, LLR from x^e (+ optional confirmed positions)."""
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 2, (F, oc.N)).astype(np.uint8)
    e = (rng.random((F, oc.N)) < qber).astype(np.uint8)
    mag = np.float32(np.log((1 - qber) / qber))
    llr = np.where(x ^ e, -mag, mag).astype(np.float32)
    if confirmed is not None:   # parity bits sent over: +-23.02585 (BOOT/src/main.cpp:19,351-354)
        c = np.float32(23.02585)
        llr[:, confirmed] = np.where(x[:, confirmed], -c, c)
    syn = np.stack([oc.syndrome(x[f]) for f in range(F)])
    return llr, syn, x


def test_kat_pegreg504x1008(q, data_dir, kat):
    """KAT-D: llrs[1008] -> decoded[504], flooding SPA ("main.cpp (alist)":443-462)."""
    code = q.Code.from_alist("%s/PEGReg504x1008.alist" % data_dir)
    assert (code.n, code.m, code.k) == (1008, 504, 504)
    for n_ite in (10, 20, 100):
        dec = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=q.RULE_SPA, dtype=q.DTYPE_F32, max_iter=n_ite)
        assert dec.kernel_name == "flooding_csr"
        out, ok, iters, _ = dec.decode(np.array([kat["llrs"]], dtype=np.float32))
        assert ok[0] and iters[0] == 6
        assert (q.unpack_bits(out, 504)[0] == np.array(kat["decoded"], dtype=np.uint8)).all()
        dec.close()


@pytest.mark.parametrize("rule", ["spa", "nms", "oms"])
@pytest.mark.parametrize("name", ["PEGReg504x1008.alist", "20.alist", "1998.5.3.2665.alist"])
def test_flooding_f32_vs_oracle(q, O, data_dir, name, rule):
    path = "%s/%s" % (data_dir, name)
    oc = O.Code.from_alist(path)
    code = q.Code.from_alist(path)
    qber = {"PEGReg504x1008.alist": 0.05, "20.alist": 0.04, "1998.5.3.2665.alist": 0.008}[name]
    F = 24
    llr, syn, x = _bsc_llr_frames(oc, F, qber, seed=3)
    qr = {"spa": q.RULE_SPA, "nms": q.RULE_NMS, "oms": q.RULE_OMS}[rule]
    orr = {"spa": O.RULE_SPA, "nms": O.RULE_NMS, "oms": O.RULE_OMS}[rule]
    norm, off = (0.8125, 0.0) if rule == "nms" else (1.0, 0.25 if rule == "oms" else 0.0)
    dec = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=qr, dtype=q.DTYPE_F32, max_iter=20, early_stop=True,
                    norm_factor=norm, offset=off, out_mode=q.OUT_ALL)
    out, ok, iters, post = dec.decode(llr, q.pack_bits(syn), want_posterior=True)
    hard, opost, oit, ook, _ = oc.batch_flooding_f32(llr, syn, rule=orr, n_ite=20, early_stop=True, norm=norm, offset=off)
    assert (q.unpack_bits(out, oc.N) == hard).all()
    assert (iters == oit).all() and (ok == ook).all()
    np.testing.assert_allclose(post, opost, rtol=RTOL, atol=1e-4)
    dec.close()


def test_flooding_send_parity_formulation(q, O, data_dir):
    """The reference's formulation: zero syndrome, parity positions carry +-23.02585, info bits decoded."""
    path = "%s/PEGReg504x1008.alist" % data_dir
    oc = O.Code.from_alist(path)
    code = q.Code.from_alist(path)
    H = oc.dense()
    # codewords: solve H_p p = H_u u with info bits in columns 504..1007 (as the reference's G does)
    rng = np.random.default_rng(9)
    F = 16
    from test_oracle import gf2_solve_parity
    cws = np.stack([gf2_solve_parity(H, rng.integers(0, 2, 504).astype(np.uint8)) for _ in range(F)])
    e = np.zeros_like(cws)
    e[:, 504:] = (rng.random((F, 504)) < 0.06).astype(np.uint8)
    mag = np.float32(np.log(0.94 / 0.06))
    llr = np.where(cws ^ e, -mag, mag).astype(np.float32)
    llr[:, :504] = np.where(cws[:, :504], -23.02585, 23.02585).astype(np.float32)
    dec = q.Decoder(code, max_iter=100)   # defaults = the reference's decoder parameters
    out, ok, iters, post = dec.decode(llr, None, want_posterior=True)
    hard, opost, oit, ook, _ = oc.batch_flooding_f32(llr, None, rule=O.RULE_SPA, n_ite=100, early_stop=True)
    assert (q.unpack_bits(out, 504) == hard[:, 504:]).all()
    assert (iters == oit).all() and (ok == ook).all()
    np.testing.assert_allclose(post, opost, rtol=RTOL, atol=1e-4)
    assert ok.all() and (q.unpack_bits(out, 504) == cws[:, 504:]).all()


@pytest.mark.parametrize("dtype", ["i8", "i16"])
@pytest.mark.parametrize("rule", ["nms", "oms"])
def test_flooding_fixed_vs_oracle(q, O, data_dir, dtype, rule):
    path = "%s/PEGReg504x1008.alist" % data_dir
    oc = O.Code.from_alist(path)
    code = q.Code.from_alist(path)
    F = 24
    rng = np.random.default_rng(4)
    x = rng.integers(0, 2, (F, oc.N)).astype(np.uint8)
    e = (rng.random((F, oc.N)) < 0.05).astype(np.uint8)
    syn = np.stack([oc.syndrome(x[f]) for f in range(F)])
    mag, vmax = (12, 127) if dtype == "i8" else (96, 32767)
    llr = np.where(x ^ e, -mag, mag)
    dt = q.DTYPE_I8 if dtype == "i8" else q.DTYPE_I16
    qr, orr = (q.RULE_NMS, O.RULE_NMS) if rule == "nms" else (q.RULE_OMS, O.RULE_OMS)
    dec = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=qr, dtype=dt, max_iter=20, early_stop=True,
                    norm_factor=0.75, offset=1.0, out_mode=q.OUT_ALL)
    out, ok, iters, post = dec.decode(llr.astype(dec.np_dtype), q.pack_bits(syn), want_posterior=True)
    for f in range(F):
        hard, opost, oit, ook = oc.decode_flooding_fixed(llr[f], syn[f], rule=orr, n_ite=20, early_stop=True, offset=1,
                                                         norm_eighths=6, vmax=vmax)
        assert (q.unpack_bits(out[f:f + 1], oc.N)[0] == hard).all()
        assert iters[f] == oit and ok[f] == ook and (post[f] == opost).all()


def test_flooding_on_qc_code_and_layered_f32(q, O, data_dir):
    """5G-NR .qc matrix through the flooding SPA decoder ("main.cpp (5g-qc)":242, n_ite=10) and the float
    layered decoder (:256-270)."""
    path = "%s/NR_1_1_24.qc" % data_dir
    oc = O.Code.from_qc(path)
    code = q.Code.from_qc_file(path)
    F = 16
    llr, syn, x = _bsc_llr_frames(oc, F, 0.03, seed=8)
    dec = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=q.RULE_SPA, dtype=q.DTYPE_F32, max_iter=10, out_mode=q.OUT_ALL)
    out, ok, iters, post = dec.decode(llr, q.pack_bits(syn), want_posterior=True)
    hard, opost, oit, ook, _ = oc.batch_flooding_f32(llr, syn, rule=O.RULE_SPA, n_ite=10, early_stop=True)
    assert (q.unpack_bits(out, oc.N) == hard).all() and (iters == oit).all() and (ok == ook).all()
    np.testing.assert_allclose(post, opost, rtol=RTOL, atol=1e-4)
    for rule, orr in ((q.RULE_SPA, O.RULE_SPA), (q.RULE_NMS, O.RULE_NMS), (q.RULE_OMS, O.RULE_OMS)):
        dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=rule, dtype=q.DTYPE_F32, max_iter=10, norm_factor=0.875,
                        offset=0.3, out_mode=q.OUT_ALL)
        assert dec.kernel_name == "layered_generic"
        out, ok, iters, post = dec.decode(llr, q.pack_bits(syn), want_posterior=True)
        for f in range(F):
            h, p, it, o = oc.decode_layered_f32(llr[f], syn[f], rule=orr, n_ite=10, early_stop=True, norm=0.875, offset=0.3)
            assert (q.unpack_bits(out[f:f + 1], oc.N)[0] == h).all() and iters[f] == it and ok[f] == o
            np.testing.assert_allclose(post[f], p, rtol=RTOL, atol=1e-4)


def test_layered_i16_vs_oracle(q, O, data_dir):
    path = "%s/NR_1_1_24.qc" % data_dir
    oc = O.Code.from_qc(path)
    code = q.Code.from_qc_file(path)
    F = 16
    rng = np.random.default_rng(2)
    x = rng.integers(0, 2, (F, oc.N)).astype(np.uint8)
    e = (rng.random((F, oc.N)) < 0.03).astype(np.uint8)
    syn = np.stack([oc.syndrome(x[f]) for f in range(F)])
    llr = np.where(x ^ e, -111, 111).astype(np.int16)
    dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_OMS, dtype=q.DTYPE_I16, max_iter=12, offset=8.0,
                    out_mode=q.OUT_ALL)
    out, ok, iters, post = dec.decode(llr, q.pack_bits(syn), want_posterior=True)
    for f in range(F):
        h, app, it, o = oc.decode_layered_fixed(llr[f].astype(np.int32), syn[f], rule=O.RULE_OMS, n_ite=12, early_stop=True,
                                                offset=8, msg_max=511, app_max=8191)
        assert (q.unpack_bits(out[f:f + 1], oc.N)[0] == h).all() and iters[f] == it and ok[f] == o
        assert (post[f] == app).all()


def test_syndrome_and_make_llr_kernels(q, O, data_dir):
    for name, loader in (("PEGReg504x1008.alist", "alist"), ("NR_2_6_52.qc", "qc"), ("test2.qc", "qc")):
        path = "%s/%s" % (data_dir, name)
        oc = O.Code.from_alist(path) if loader == "alist" else O.Code.from_qc(path)
        code = q.Code.from_alist(path) if loader == "alist" else q.Code.from_qc_file(path)
        dec = q.Decoder(code, max_iter=1)
        rng = np.random.default_rng(1)
        F = 9
        bits = rng.integers(0, 2, (F, oc.N)).astype(np.uint8)
        syn = q.unpack_bits(dec.syndrome(q.pack_bits(bits)), oc.M)
        for f in range(F):
            assert (syn[f] == oc.syndrome(bits[f])).all()
        known = rng.integers(0, 2, oc.N).astype(np.uint8)
        punct = (rng.random(oc.N) < 0.1).astype(np.uint8)
        llr = dec.make_llr(q.pack_bits(bits), 3.4761, 23.02585, q.pack_bits(known), q.pack_bits(punct))
        mag = np.where(punct == 1, 0.0, np.where(known == 1, 23.02585, 3.4761)).astype(np.float32)
        np.testing.assert_array_equal(llr, np.where(bits == 1, -mag, mag).astype(np.float32))


def test_config1_n1944_spa_flooding_vs_oracle(q, O, data_dir):
    """BASELINE config 1 on the GPU path: same frames as the CPU plumbing test, bit-exact bits / iterations / ok"""
    path = "%s/wifi_n1944_r12.qc" % data_dir
    oc = O.Code.from_qc(path)
    code = q.Code.from_qc_file(path)
    rng = np.random.default_rng(1944)
    F, qb = 64, 0.03
    x = rng.integers(0, 2, (F, oc.N)).astype(np.uint8)
    e = (rng.random((F, oc.N)) < qb).astype(np.uint8)
    syn = np.stack([oc.syndrome(x[f]) for f in range(F)])
    mag = np.float32(np.log((1 - qb) / qb))
    llr = np.where(x ^ e, -mag, mag).astype(np.float32)
    dec = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=q.RULE_SPA, dtype=q.DTYPE_F32, max_iter=20, early_stop=True,
                    out_mode=q.OUT_ALL)
    out, ok, iters, post = dec.decode(llr, q.pack_bits(syn), want_posterior=True)
    hard, opost, oit, ook, _ = oc.batch_flooding_f32(llr, syn, rule=O.RULE_SPA, n_ite=20, early_stop=True)
    assert (q.unpack_bits(out, oc.N) == hard).all() and (hard == x).all()
    assert (iters == oit).all() and (ok == ook).all() and ok.all()
    np.testing.assert_allclose(post, opost, rtol=RTOL, atol=1e-4)
    dec.close()


@pytest.mark.parametrize("name,qber", [("PEGReg504x1008.alist", 0.05), ("wifi_n1944_r12.qc", 0.03), ("1998.5.3.2665.alist", 0.008)])
def test_fast_spa_flavour_on_the_general_flooding_kernel(q, O, data_dir, kat, name, qber):
    """QLDPC_FLAG_FAST_SPA (fp32 tanh / atanh on the special-function units) on flooding_csr: the reference's KAT decodes to
    the same word after the same 6 sweeps; on random batches the decoded bits, flags and iteration counts equal the exact
    flavour's (= the oracle's), posteriors within 1e-3 except for saturated messages"""
    path = "%s/%s" % (data_dir, name)
    alist = name.endswith(".alist")
    oc = O.Code.from_alist(path) if alist else O.Code.from_qc(path)
    code = q.Code.from_alist(path) if alist else q.Code.from_qc_file(path)
    fast = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=q.RULE_SPA, dtype=q.DTYPE_F32, max_iter=20, early_stop=True,
                     out_mode=q.OUT_ALL, flags=q.FLAG_FAST_SPA)
    if name.startswith("PEGReg"):
        out, ok, iters, _ = fast.decode(np.array([kat["llrs"]], dtype=np.float32))
        assert ok[0] and iters[0] == 6
        assert (q.unpack_bits(out, oc.N)[0][oc.N - 504:] == np.array(kat["decoded"], dtype=np.uint8)).all()
    F = 64
    llr, syn, x = _bsc_llr_frames(oc, F, qber, seed=11)
    out, ok, iters, post = fast.decode(llr, q.pack_bits(syn), want_posterior=True)
    hard, opost, oit, ook, _ = oc.batch_flooding_f32(llr, syn, rule=O.RULE_SPA, n_ite=20, early_stop=True)
    assert (q.unpack_bits(out, oc.N) == hard).all() and (iters == oit).all() and (ok == ook).all()
    dev = np.abs(post - opost)
    outside = dev > 1e-3 * np.abs(opost) + 1e-4
    # what falls outside 1e-3: saturated messages (one last-bit difference in 1 - r), or a posterior that is the small
    # difference of large terms (absolute deviation still below 1e-2)
    assert outside.mean() < 2e-3 and dev.max() < 1.5 and ((np.abs(opost[outside]) > 15) | (dev[outside] < 1e-2)).all()
    fast.close()


@pytest.mark.parametrize("rule", ["nms", "oms", "spa"])
def test_layered_f32_bg1_z384_compressed_check_state(q, O, data_dir, rule):
    """float layered decoding on BG1 Z=384 (rows of 19 edges): the min-sum rules keep {c1, c2, index, signs} per check lane
    instead of one message per edge (the per-edge messages of the frames in flight would not fit in L2), SPA keeps per-edge
    messages -- both against the oracle, posteriors included"""
    path = "%s/NR_1_1_384.qc" % data_dir
    oc = O.Code.from_qc(path)
    code = q.Code.from_qc_file(path)
    F = 3
    llr, syn, x = _bsc_llr_frames(oc, F, 0.06, seed=21)
    qr = {"spa": q.RULE_SPA, "nms": q.RULE_NMS, "oms": q.RULE_OMS}[rule]
    orr = {"spa": O.RULE_SPA, "nms": O.RULE_NMS, "oms": O.RULE_OMS}[rule]
    dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=qr, dtype=q.DTYPE_F32, max_iter=8, norm_factor=0.8125, offset=0.4,
                    out_mode=q.OUT_ALL)
    assert dec.kernel_name == "layered_generic"
    out, ok, iters, post = dec.decode(llr, q.pack_bits(syn), want_posterior=True)
    for f in range(F):
        h, p, it, o = oc.decode_layered_f32(llr[f], syn[f], rule=orr, n_ite=8, early_stop=True, norm=0.8125, offset=0.4)
        assert (q.unpack_bits(out[f:f + 1], oc.N)[0] == h).all() and iters[f] == it and ok[f] == o
        np.testing.assert_allclose(post[f], p, rtol=RTOL, atol=1e-4)
    assert ok.all() and (q.unpack_bits(out, oc.N) == x).all()
    dec.close()


@pytest.mark.parametrize("name", ["NR_1_1_24.qc", "PEGReg504x1008.alist"])
def test_fast_spa_flavour_on_the_layered_kernels(q, O, data_dir, name):
    """QLDPC_FLAG_FAST_SPA on the float layered kernels (layered_generic on a QC code, layered_csr on an .alist H): decoded bits,
    flags and iteration counts equal the oracle's, posteriors within 1e-3 outside saturated / cancelling entries"""
    path = "%s/%s" % (data_dir, name)
    alist = name.endswith(".alist")
    oc = O.Code.from_alist(path) if alist else O.Code.from_qc(path)
    code = q.Code.from_alist(path) if alist else q.Code.from_qc_file(path)
    F = 24
    llr, syn, x = _bsc_llr_frames(oc, F, 0.04, seed=13)
    dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_SPA, dtype=q.DTYPE_F32, max_iter=12, out_mode=q.OUT_ALL,
                    flags=q.FLAG_FAST_SPA)
    assert dec.kernel_name == ("layered_csr" if alist else "layered_generic")
    out, ok, iters, post = dec.decode(llr, q.pack_bits(syn), want_posterior=True)
    for f in range(F):
        h, p, it, o = oc.decode_layered_f32(llr[f], syn[f], rule=O.RULE_SPA, n_ite=12, early_stop=True)
        assert (q.unpack_bits(out[f:f + 1], oc.N)[0] == h).all() and iters[f] == it and ok[f] == o
        dev = np.abs(post[f] - p)
        outside = dev > 1e-3 * np.abs(p) + 1e-4
        assert outside.mean() < 5e-3 and ((np.abs(p[outside]) > 15) | (dev[outside] < 1e-2)).all()
    dec.close()
