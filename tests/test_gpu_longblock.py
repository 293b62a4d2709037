"""BASELINE config 3: QKD long-block QC-LDPC (N = 65536, reference PSD-PEG construction, Z = 2048) with rate adaptation
by puncturing / shortening, float flooding SPA vs min-sum -- CUDA path (through the C ABI) against the CPU oracle.
Bar: decoded bits, syndrome-ok flag and iteration count bit-exact; posteriors within 1e-3 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CODE = "qkd_psdpeg_n65536.qc"


def _frames(q, oc, F, qber, punct_frac, short_frac, seed):
    """syndrome formulation with modulation: Alice's word x (punctured positions: random filler, shortened positions:
    publicly known), Bob's word y = x ^ e on the key positions.  Returns packed bits for make_llr and the masks."""
    rng = np.random.default_rng(seed)
    N = oc.N
    perm = np.random.default_rng(7).permutation(N)            # the modulation pattern is public and fixed
    n_p, n_s = int(punct_frac * N), int(short_frac * N)
    punct = np.zeros(N, np.uint8); punct[perm[:n_p]] = 1
    short = np.zeros(N, np.uint8); short[perm[n_p:n_p + n_s]] = 1
    x = rng.integers(0, 2, (F, N)).astype(np.uint8)
    e = (rng.random((F, N)) < qber).astype(np.uint8)
    e[:, (punct | short) == 1] = 0
    y = x ^ e
    syn = np.stack([oc.syndrome(x[f]) for f in range(F)])
    return x, y, syn, punct, short


@pytest.mark.parametrize("rule,norm", [("spa", 1.0), ("nms", 0.8125)])
@pytest.mark.parametrize("qber,pf,sf", [(0.06, 0.0, 0.0), (0.05, 0.10, 0.02), (0.03, 0.20, 0.0)])
def test_n65536_flooding_rate_adapted_vs_oracle(q, O, data_dir, rule, norm, qber, pf, sf):
    path = "%s/%s" % (data_dir, CODE)
    oc = O.Code.from_qc(path)
    assert (oc.N, oc.M) == (65536, 32768)
    code = q.Code.from_qc_file(path)
    qr = q.RULE_SPA if rule == "spa" else q.RULE_NMS
    orr = O.RULE_SPA if rule == "spa" else O.RULE_NMS
    dec = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=qr, dtype=q.DTYPE_F32, max_iter=30, early_stop=True,
                    norm_factor=norm, out_mode=q.OUT_ALL)
    assert dec.kernel_name == "flooding_qc_cluster"        # one frame per thread-block cluster, state in L2 (flooding_qcx.cu)
    F = 3
    x, y, syn, punct, short = _frames(q, oc, F, qber, pf, sf, seed=int(qber * 1000) + 1)
    mag = float(np.log((1 - qber) / qber))
    # shortened positions carry Alice's (public) bit with the reference's "confirmed" magnitude (BOOT/src/main.cpp:19)
    llr = dec.make_llr(q.pack_bits(y), mag, 23.02585, known_mask=q.pack_bits(short[None, :])[0],
                       punct_mask=q.pack_bits(punct[None, :])[0])
    assert (llr[:, punct == 1] == 0).all() and (np.abs(llr[:, short == 1]) > 23).all()
    out, ok, iters, post = dec.decode(llr, q.pack_bits(syn), want_posterior=True)
    hard, opost, oit, ook, _ = oc.batch_flooding_f32(llr, syn, rule=orr, n_ite=30, early_stop=True, norm=norm, offset=0.0)
    got = q.unpack_bits(out, oc.N)
    assert (got == hard).all()
    assert (iters == oit).all() and (ok == ook).all()
    np.testing.assert_allclose(post, opost, rtol=1e-3, atol=1e-4)
    if ok.all():   # converged frames reproduce Alice's word, filler bits included
        assert (got == x).all()
    dec.close()
    if rule == "spa":
        # QLDPC_FLAG_FAST_SPA: fp32 transcendentals.  Same bits, flags and iteration counts; posteriors within 1e-3 except for
        # saturated messages, where one last-bit difference in 1 - r moves 2 atanh(r) by ln 2 / ln 3/2 / ...
        fast = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=qr, dtype=q.DTYPE_F32, max_iter=30, early_stop=True,
                         out_mode=q.OUT_ALL, flags=q.FLAG_FAST_SPA)
        out2, ok2, iters2, post2 = fast.decode(llr, q.pack_bits(syn), want_posterior=True)
        assert (out2 == out).all() and (ok2 == ok).all() and (iters2 == iters).all()
        dev = np.abs(post2 - opost)
        outside = dev > 1e-3 * np.abs(opost) + 1e-4
        assert outside.mean() < 1e-3 and dev.max() < 1.5 and (np.abs(opost[outside]) > 20).all()
        fast.close()


@pytest.mark.parametrize("rule", ["spa", "nms"])
def test_n65536_more_frames_than_clusters(q, O, data_dir, rule):
    """more frames than clusters are co-resident (37 clusters of 4 SMs on a B200): every cluster switches frames and reuses
    its scratch; the first sweep of a frame reads no stale messages; a mix of iteration counts inside the batch"""
    path = "%s/%s" % (data_dir, CODE)
    oc = O.Code.from_qc(path)
    code = q.Code.from_qc_file(path)
    qr, orr, norm = (q.RULE_SPA, O.RULE_SPA, 1.0) if rule == "spa" else (q.RULE_NMS, O.RULE_NMS, 0.8125)
    dec = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=qr, dtype=q.DTYPE_F32, max_iter=12, early_stop=True,
                    norm_factor=norm, out_mode=q.OUT_ALL)
    F = 90
    rng = np.random.default_rng(65)
    x = rng.integers(0, 2, (F, oc.N)).astype(np.uint8)
    qbers = np.linspace(0.01, 0.075, F)                         # the last frames do not converge in 12 sweeps
    y = x ^ (rng.random((F, oc.N)) < qbers[:, None])
    syn_p = dec.syndrome(q.pack_bits(x))
    llr = np.where(y, -2.9, 2.9).astype(np.float32)
    out, ok, iters, _ = dec.decode(llr, syn_p)
    hard, _, oit, ook, _ = oc.batch_flooding_f32(llr, q.unpack_bits(syn_p, oc.M), rule=orr, n_ite=12, early_stop=True, norm=norm)
    assert (iters == oit).all() and (ok == ook).all() and (q.unpack_bits(out, oc.N) == hard).all()
    assert ok[:40].all() and not ok.all() and len(set(iters.tolist())) >= 4
    dec.close()


@pytest.mark.parametrize("dtype,rule", [("i8", "nms"), ("i8", "oms"), ("i16", "nms")])
def test_n65536_fixed_point_flooding_vs_oracle(q, O, data_dir, dtype, rule):
    """int8 / int16 flooding min-sum on the long block (messages stored in 8 / 16 bits): bits, iteration counts and integer
    posteriors equal to the oracle's"""
    path = "%s/%s" % (data_dir, CODE)
    oc = O.Code.from_qc(path)
    code = q.Code.from_qc_file(path)
    dt, mag, vmax = (q.DTYPE_I8, 12, 127) if dtype == "i8" else (q.DTYPE_I16, 200, 32767)
    qr, orr = (q.RULE_NMS, O.RULE_NMS) if rule == "nms" else (q.RULE_OMS, O.RULE_OMS)
    dec = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=qr, dtype=dt, max_iter=20, early_stop=True, norm_factor=0.75,
                    offset=1.0, out_mode=q.OUT_ALL)
    assert dec.kernel_name == "flooding_qc_cluster"
    F = 6
    rng = np.random.default_rng(3)
    x = rng.integers(0, 2, (F, oc.N)).astype(np.uint8)
    y = x ^ (rng.random((F, oc.N)) < 0.04)
    syn_p = dec.syndrome(q.pack_bits(x))
    syn = q.unpack_bits(syn_p, oc.M)
    llr = np.where(y, -mag, mag)
    out, ok, iters, post = dec.decode(llr.astype(dec.np_dtype), syn_p, want_posterior=True)
    for f in range(F):
        hard, opost, oit, ook = oc.decode_flooding_fixed(llr[f], syn[f], rule=orr, n_ite=20, early_stop=True, offset=1,
                                                         norm_eighths=6, vmax=vmax)
        assert (q.unpack_bits(out[f:f + 1], oc.N)[0] == hard).all() and iters[f] == oit and ok[f] == ook
        assert (post[f] == opost).all()
    assert ok.all()
    dec.close()


IRREGULAR = "qkd_irregular_n65536_r34.qc"


@pytest.mark.parametrize("tier", ["spa", "nms", "i8", "i16"])
def test_n65536_irregular_rate34_flooding_vs_oracle(q, O, data_dir, tier):
    """the irregular rate-3/4 code (Z = 1024, row degrees 17 / 18): heavy rows take the 2-lanes-per-thread instantiation of the
    clustered kernel, whose per-degree code paths the (3,6) code never reaches"""
    path = "%s/%s" % (data_dir, IRREGULAR)
    oc = O.Code.from_qc(path)
    assert (oc.N, oc.M) == (65536, 16384)
    code = q.Code.from_qc_file(path)
    F, qber, n_ite = 3, 0.025, 25
    rng = np.random.default_rng(34)
    x = rng.integers(0, 2, (F, oc.N)).astype(np.uint8)
    y = x ^ (rng.random((F, oc.N)) < qber)
    if tier in ("spa", "nms"):
        qr, orr, norm = (q.RULE_SPA, O.RULE_SPA, 1.0) if tier == "spa" else (q.RULE_NMS, O.RULE_NMS, 0.8125)
        dec = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=qr, dtype=q.DTYPE_F32, max_iter=n_ite, early_stop=True,
                        norm_factor=norm, out_mode=q.OUT_ALL)
        assert dec.kernel_name == "flooding_qc_cluster"
        syn_p = dec.syndrome(q.pack_bits(x))
        mag = float(np.log((1 - qber) / qber))
        llr = np.where(y, -mag, mag).astype(np.float32)
        out, ok, iters, post = dec.decode(llr, syn_p, want_posterior=True)
        hard, opost, oit, ook, _ = oc.batch_flooding_f32(llr, q.unpack_bits(syn_p, oc.M), rule=orr, n_ite=n_ite, early_stop=True,
                                                         norm=norm)
        assert (q.unpack_bits(out, oc.N) == hard).all() and (iters == oit).all() and (ok == ook).all()
        np.testing.assert_allclose(post, opost, rtol=1e-3, atol=1e-4)
    else:
        dt, mag, vmax = (q.DTYPE_I8, 14, 127) if tier == "i8" else (q.DTYPE_I16, 230, 32767)
        dec = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=q.RULE_NMS, dtype=dt, max_iter=n_ite, early_stop=True,
                        norm_factor=0.75, out_mode=q.OUT_ALL)
        assert dec.kernel_name == "flooding_qc_cluster"
        syn_p = dec.syndrome(q.pack_bits(x))
        syn = q.unpack_bits(syn_p, oc.M)
        llr = np.where(y, -mag, mag)
        out, ok, iters, post = dec.decode(llr.astype(dec.np_dtype), syn_p, want_posterior=True)
        for f in range(F):
            hard, opost, oit, ook = oc.decode_flooding_fixed(llr[f], syn[f], rule=O.RULE_NMS, n_ite=n_ite, early_stop=True,
                                                             norm_eighths=6, vmax=vmax)
            assert (q.unpack_bits(out[f:f + 1], oc.N)[0] == hard).all() and iters[f] == oit and ok[f] == ook
            assert (post[f] == opost).all()
    assert ok.all() and (q.unpack_bits(out, oc.N) == x).all()
    dec.close()


@pytest.mark.parametrize("tier", ["f32", "i16"])
def test_n65536_layered_schedule_on_the_long_block(q, O, data_dir, tier):
    """the generic layered kernel outside its comfortable sizes: Z = 2048 lanes on 1024 threads, and (float) a frame whose
    beliefs (262 KB) do not fit in shared memory and live in the global scratch instead"""
    path = "%s/%s" % (data_dir, CODE)
    oc = O.Code.from_qc(path)
    code = q.Code.from_qc_file(path)
    F = 2
    rng = np.random.default_rng(5)
    x = rng.integers(0, 2, (F, oc.N)).astype(np.uint8)
    y = x ^ (rng.random((F, oc.N)) < 0.04)
    syn = np.stack([oc.syndrome(x[f]) for f in range(F)])
    if tier == "f32":
        dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_F32, max_iter=12, norm_factor=0.8125,
                        out_mode=q.OUT_ALL)
        llr = np.where(y, -3.1, 3.1).astype(np.float32)
    else:
        dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_OMS, dtype=q.DTYPE_I16, max_iter=12, offset=8.0,
                        out_mode=q.OUT_ALL)
        llr = np.where(y, -111, 111).astype(np.int16)
    assert dec.kernel_name == "layered_generic"
    out, ok, iters, post = dec.decode(llr, q.pack_bits(syn), want_posterior=True)
    for f in range(F):
        if tier == "f32":
            h, p, it, o = oc.decode_layered_f32(llr[f], syn[f], rule=O.RULE_NMS, n_ite=12, early_stop=True, norm=0.8125)
            np.testing.assert_allclose(post[f], p, rtol=1e-3, atol=1e-4)
        else:
            h, p, it, o = oc.decode_layered_fixed(llr[f].astype(np.int32), syn[f], rule=O.RULE_OMS, n_ite=12, early_stop=True,
                                                  offset=8, msg_max=511, app_max=8191)
            assert (post[f] == p).all()
        assert (q.unpack_bits(out[f:f + 1], oc.N)[0] == h).all() and iters[f] == it and ok[f] == o
    assert ok.all()
    dec.close()


@pytest.mark.parametrize("tier", ["nms", "spa", "i8"])
def test_clustered_flooding_rows_heavier_than_the_compiled_degrees(q, O, tier):
    """a synthetic QC code with block rows of 24 edges (the clustered kernel compiles row degrees up to 20): heavy rows take
    the two-pass loop, the light row a compiled variant; messages + posteriors (295 KB) do not fit in shared memory"""
    rng = np.random.default_rng(24)
    rows, cols, Z = 10, 48, 256
    base = -np.ones((rows, cols), np.int32)
    for r in range(rows):
        deg = 6 if r == 0 else 24
        pick = rng.choice(cols, deg, replace=False)
        base[r, pick] = rng.integers(0, Z, deg)
    for c in range(cols):                                     # no empty block column
        if (base[:, c] < 0).all():
            base[rng.integers(1, rows), c] = rng.integers(0, Z)
    oc = O.Code.from_base(base, Z)
    code = q.Code.from_qc(base, Z)
    F = 5
    x = rng.integers(0, 2, (F, oc.N)).astype(np.uint8)
    y = x ^ (rng.random((F, oc.N)) < 0.004)
    syn = np.stack([oc.syndrome(x[f]) for f in range(F)])
    if tier == "i8":
        dec = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=15, early_stop=True,
                        norm_factor=0.75, out_mode=q.OUT_ALL)
        assert dec.kernel_name == "flooding_qc_cluster"
        llr = np.where(y, -20, 20)
        out, ok, iters, post = dec.decode(llr.astype(np.int8), q.pack_bits(syn), want_posterior=True)
        for f in range(F):
            hard, opost, oit, ook = oc.decode_flooding_fixed(llr[f], syn[f], rule=O.RULE_NMS, n_ite=15, early_stop=True,
                                                             norm_eighths=6, vmax=127)
            assert (q.unpack_bits(out[f:f + 1], oc.N)[0] == hard).all() and iters[f] == oit and ok[f] == ook
            assert (post[f] == opost).all()
    else:
        qr, orr, norm = (q.RULE_SPA, O.RULE_SPA, 1.0) if tier == "spa" else (q.RULE_NMS, O.RULE_NMS, 0.75)
        dec = q.Decoder(code, schedule=q.SCHED_FLOODING, rule=qr, dtype=q.DTYPE_F32, max_iter=15, early_stop=True,
                        norm_factor=norm, out_mode=q.OUT_ALL)
        assert dec.kernel_name == "flooding_qc_cluster"
        llr = np.where(y, -5.5, 5.5).astype(np.float32)
        out, ok, iters, post = dec.decode(llr, q.pack_bits(syn), want_posterior=True)
        hard, opost, oit, ook, _ = oc.batch_flooding_f32(llr, syn, rule=orr, n_ite=15, early_stop=True, norm=norm)
        assert (q.unpack_bits(out, oc.N) == hard).all() and (iters == oit).all() and (ok == ook).all()
        np.testing.assert_allclose(post, opost, rtol=1e-3, atol=1e-4)
    dec.close()
