"""The CPU oracle against every golden vector / structural fact the reference holds for the LDPC path
(SURVEY.md 8c).  No GPU."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gf2_solve_parity(H, info, info_cols=None):
    """systematic encoding from H: codeword c with c[info_cols] = info and H c = 0 (parity columns are the rest)"""
    M, N = H.shape
    if info_cols is None:
        info_cols = np.arange(N - (N - M), N) if False else np.arange(M, N)
    par_cols = np.setdiff1d(np.arange(N), info_cols)
    A = H[:, par_cols].astype(np.uint8).copy()
    b = (H[:, info_cols].astype(np.int64) @ info.astype(np.int64) % 2).astype(np.uint8)
    n = A.shape[1]
    Ab = np.concatenate([A, b[:, None]], axis=1)
    row = 0
    piv = []
    for col in range(n):
        r = row + np.nonzero(Ab[row:, col])[0]
        if r.size == 0:
            continue
        Ab[[row, r[0]]] = Ab[[r[0], row]]
        others = np.nonzero(Ab[:, col])[0]
        others = others[others != row]
        Ab[others] ^= Ab[row]
        piv.append(col)
        row += 1
    assert len(piv) == n, "parity part of H is singular"
    p = np.zeros(n, dtype=np.uint8)
    p[piv] = Ab[:len(piv), -1]
    c = np.zeros(N, dtype=np.uint8)
    c[info_cols] = info
    c[par_cols] = p
    return c


def test_kat_encoder(O, data_dir, kat):
    """KAT-E: data[504] -> encoded[1008]; H*encoded = 0, encoded[504:] == data, and the systematic
    encoding derived from H alone reproduces the reference's codeword."""
    c = O.Code.from_alist("%s/PEGReg504x1008.alist" % data_dir)
    assert (c.N, c.M, c.E) == (1008, 504, 3024)
    enc = np.array(kat["encoded"], dtype=np.uint8)
    data = np.array(kat["data"], dtype=np.uint8)
    assert not c.syndrome(enc).any()
    assert (enc[504:] == data).all()
    assert (gf2_solve_parity(c.dense(), data) == enc).all()


@pytest.mark.parametrize("n_ite", [10, 20, 100])
def test_kat_decoder(O, data_dir, kat, n_ite):
    """KAT-D: llrs[1008] -> decoded[504], flooding SPA, info bits at 504..1007."""
    c = O.Code.from_alist("%s/PEGReg504x1008.alist" % data_dir)
    llr = np.array(kat["llrs"], dtype=np.float32)
    assert int(c.syndrome((llr < 0).astype(np.uint8)).sum()) == 147
    hard, post, it, ok = c.decode_flooding_f32(llr, n_ite=n_ite)
    assert ok and it == 6
    assert (hard[504:] == np.array(kat["decoded"], dtype=np.uint8)).all()
    assert int(((llr < 0)[504:] != hard[504:]).sum()) == 32               # fixes 32 info-bit errors


def test_kat_decoder_other_rules_agree_on_bits(O, data_dir, kat):
    c = O.Code.from_alist("%s/PEGReg504x1008.alist" % data_dir)
    llr = np.array(kat["llrs"], dtype=np.float32)
    want = np.array(kat["decoded"], dtype=np.uint8)
    for rule, norm, off in ((O.RULE_NMS, 0.8125, 0.0), (O.RULE_OMS, 1.0, 0.3)):
        hard, _, _, ok = c.decode_flooding_f32(llr, rule=rule, n_ite=50, norm=norm, offset=off)
        assert ok and (hard[504:] == want).all()
    hard, _, _, ok = c.decode_layered_f32(llr, rule=O.RULE_SPA, n_ite=50)
    assert ok and (hard[504:] == want).all()


def test_matrix_facts(O, data_dir):
    facts = json.load(open(os.path.join(ROOT, "tests", "golden", "matrix_facts.json")))
    for fn, (rows, cols, edges) in facts["all_nr_files"].items():
        if fn.startswith("NR_1_"):
            assert (rows, cols, edges) == (46, 68, 316), fn
        else:
            assert (rows, cols, edges) == (42, 52, 197), fn
    c = O.Code.from_qc("%s/NR_1_1_384.qc" % data_dir)
    assert (c.N, c.M, c.K, c.E, c.Z) == (26112, 17664, 8448, 121344, 384)
    b = c.base
    assert [(r >= 0).sum() for r in b] == facts["NR_1_1_384"]["row_degrees"]
    assert [(r >= 0).sum() for r in b][:10] == [19, 19, 19, 19, 3, 8, 9, 7, 10, 9]
    assert [(col >= 0).sum() for col in b.T][:6] == [30, 28, 7, 11, 9, 4]
    # SURVEY 8c pin 3: NR_1_1_384 mod 192 == NR_1_1_192.qc etc.
    for big, small, z in (("NR_1_1_384.qc", "NR_1_1_192.qc", 192), ("NR_1_0_256.qc", "NR_1_0_2.qc", 2),
                          ("NR_1_7_240.qc", "NR_1_7_30.qc", 30)):
        bb = O.Code.from_qc("%s/%s" % (data_dir, big)).base
        bs = O.Code.from_qc("%s/%s" % (data_dir, small)).base
        assert (np.where(bb >= 0, bb % z, -1) == bs).all()


def test_alist_and_qc_parsers(O, data_dir):
    c = O.Code.from_alist("%s/20.alist" % data_dir)
    assert (c.N, c.M) == (504, 252)
    H = c.dense()
    assert H.sum(axis=0).min() == 2 and H.sum(axis=0).max() == 20        # irregular, dv 2..20
    c = O.Code.from_alist("%s/1998.5.3.2665.alist" % data_dir)
    H = c.dense()
    assert (c.N, c.M) == (1998, 222) and (H.sum(axis=0) == 4).all() and (H.sum(axis=1) == 36).all()
    c = O.Code.from_qc("%s/test2.qc" % data_dir)                          # shifts >= Z are reduced mod Z
    assert c.Z == 7 and c.base.max() < 7 and (c.N, c.M) == (18 * 7, 6 * 7)
    c = O.Code.from_qc("%s/test.qc" % data_dir)
    assert c.Z == 253 and (c.bcols, c.brows) == (64, 14)


def test_circulant_convention(O):
    """check lane i of block row r touches variable lane (i + shift) mod z (ML/mul_sh.m:9)."""
    base = np.array([[2, -1, 0], [-1, 1, 3]], dtype=np.int32)
    c = O.Code.from_base(base, 5)
    H = c.dense()
    for r in range(2):
        for col in range(3):
            s = base[r, col]
            blk = H[r * 5:(r + 1) * 5, col * 5:(col + 1) * 5]
            if s < 0:
                assert not blk.any()
            else:
                for i in range(5):
                    assert blk[i].sum() == 1 and blk[i, (i + s) % 5] == 1


@pytest.mark.parametrize("name", ["NR_1_1_384.qc", "NR_1_1_24.qc", "NR_2_6_52.qc", "NR_2_3_112.qc", "NR_1_0_2.qc"])
def test_nr_encoder(O, data_dir, name):
    c = O.Code.from_qc("%s/%s" % (data_dir, name))
    rng = np.random.default_rng(0)
    for _ in range(3):
        msg = rng.integers(0, 2, c.K).astype(np.uint8)
        cw = c.nr_encode(msg)
        assert (cw[:c.K] == msg).all() and not c.syndrome(cw).any()


def test_wifi_table_is_a_valid_code(O, data_dir):
    c = O.Code.from_qc("%s/wifi_n1944_r12.qc" % data_dir)
    assert (c.N, c.M, c.E, c.Z) == (1944, 972, 6966, 81)


def test_normalize_eighths(O):
    for v in range(0, 40):
        assert O.normalize_eighths(v, 8) == v
        assert O.normalize_eighths(v, 6) == (v >> 1) + (v >> 2)
        assert O.normalize_eighths(v, 4) == v >> 1
        assert O.normalize_eighths(v, 7) == (v >> 1) + (v >> 2) + (v >> 3)


def _numpy_matlab_fp(base, Z, llr, n_ite, offset, msg_max=31, app_max=127):
    """independent vectorised transcription of ML/BPSK_nrldpc_sim_FP.m:40-94 (second opinion on the C oracle)"""
    mb, nb = base.shape
    L = llr.astype(np.int64).copy()
    R = {}
    for _ in range(n_ite):
        for lyr in range(mb):
            cols = [c for c in range(nb) if base[lyr, c] >= 0]
            treg = []
            for c in cols:
                Lc = L[c * Z:(c + 1) * Z] - R.get((lyr, c), 0)
                L[c * Z:(c + 1) * Z] = Lc
                treg.append(np.clip(np.roll(Lc, -base[lyr, c]), -(msg_max + 1), msg_max))
            T = np.stack(treg)
            A = np.abs(T)
            pos = A.argmin(axis=0)
            min1 = A.min(axis=0)
            A2 = A.copy()
            A2[pos, np.arange(Z)] = 1 << 20
            min2 = A2.min(axis=0)
            S = np.where(T >= 0, 1, -1)
            par = S.prod(axis=0)
            min1 = np.maximum(min1 - offset, 0)
            min2 = np.maximum(min2 - offset, 0)
            out = np.tile(min1, (len(cols), 1))
            out[pos, np.arange(Z)] = min2
            out = par * S * out
            for j, c in enumerate(cols):
                Rn = np.roll(out[j], base[lyr, c])
                R[(lyr, c)] = Rn
                L[c * Z:(c + 1) * Z] = np.clip(L[c * Z:(c + 1) * Z] + Rn, -(app_max + 1), app_max)
    return (L < 0).astype(np.uint8), L


@pytest.mark.parametrize("name", ["NR_1_1_24.qc", "NR_2_6_52.qc"])
def test_layered_fixed_matches_independent_transcription(O, data_dir, name):
    """The two base graphs the MATLAB scripts use (BPSK_nrldpc_sim_FP.m:8, _RM_FP.m:8), with the .m
    quantiser (:35-37) on an AWGN frame: C oracle == numpy transcription, bit for bit."""
    c = O.Code.from_qc("%s/%s" % (data_dir, name))
    rng = np.random.default_rng(3)
    for trial in range(3):
        msg = rng.integers(0, 2, c.K).astype(np.uint8)
        cw = c.nr_encode(msg)
        r = (1.0 - 2.0 * cw) + 0.8 * rng.standard_normal(c.N)
        r[:2 * c.Z] = 0                                     # :33 puncturing of the first two block columns
        rq = np.clip(np.floor(r / 4 * 31), -32, 31).astype(np.int32)
        hard, app, it, ok = c.decode_layered_fixed(rq, None, rule=O.RULE_OMS, n_ite=20, early_stop=False, offset=2)
        h2, L2 = _numpy_matlab_fp(c.base, c.Z, rq, 20, 2)
        assert it == 20 and (hard == h2).all() and (app == L2).all()


def test_layered_fixed_extensions(O, data_dir):
    """syndrome folding and early stop: decoding x^e against syndrome(x) equals decoding the all-zero
    coset shifted by x (linearity), and early stop never changes the bits of a converged frame."""
    c = O.Code.from_qc("%s/NR_1_1_24.qc" % data_dir)
    rng = np.random.default_rng(5)
    x = rng.integers(0, 2, c.N).astype(np.uint8)
    e = (rng.random(c.N) < 0.03).astype(np.uint8)
    llr_x = np.where(x ^ e, -14, 14).astype(np.int32)
    llr_0 = np.where(e, -14, 14).astype(np.int32)
    for rule, kw in ((O.RULE_NMS, dict(norm_eighths=6)), (O.RULE_OMS, dict(offset=2))):
        hx, _, itx, okx = c.decode_layered_fixed(llr_x, c.syndrome(x), rule=rule, n_ite=20, early_stop=True, **kw)
        h0, _, it0, ok0 = c.decode_layered_fixed(llr_0, None, rule=rule, n_ite=20, early_stop=True, **kw)
        assert okx and ok0 and itx == it0 and ((hx ^ x) == h0).all() and not h0.any()
        hf, _, itf, okf = c.decode_layered_fixed(llr_x, c.syndrome(x), rule=rule, n_ite=20, early_stop=False, **kw)
        assert itf == 20 and okf and (hf == hx).all()


def test_batch_helpers_match_single_frame(O, data_dir):
    c = O.Code.from_qc("%s/NR_1_1_24.qc" % data_dir)
    rng = np.random.default_rng(6)
    F = 7
    llr = rng.integers(-40, 40, (F, c.N)).astype(np.int8)
    syn = rng.integers(0, 2, (F, c.M)).astype(np.uint8)
    hard, iters, ok, nt = c.batch_layered_fixed_i8(llr, syn, rule=O.RULE_NMS, n_ite=6, early_stop=True, norm_eighths=6,
                                                   n_threads=3)
    assert nt == 3
    for f in range(F):
        h, _, it, o = c.decode_layered_fixed(llr[f].astype(np.int32), syn[f], rule=O.RULE_NMS, n_ite=6, early_stop=True,
                                             norm_eighths=6)
        assert (h == hard[f]).all() and it == iters[f] and o == ok[f]


def test_pack_bits_msb_first(O):
    bits = np.zeros(40, np.uint8)
    bits[0] = 1
    bits[33] = 1
    w = O.pack_bits_msb(bits)
    assert w.tolist() == [0x80000000, 0x40000000]                         # helpers.h:68 uint32AllZeroExceptAtN
    assert (O.unpack_bits_msb(w, 40) == bits).all()


def test_config1_n1944_spa_flooding_syndrome_decoding_on_cpu(O, data_dir):
    """BASELINE config 1 (plumbing, no GPU): rate-1/2 N=1944 code, float SPA flooding, 20 iterations, BSC QBER 3 %,
    syndrome decoding as the ldpc_examples chain does it (LLR = +-ln((1-q)/q), early stop on the syndrome)."""
    oc = O.Code.from_qc("%s/wifi_n1944_r12.qc" % data_dir)
    assert (oc.N, oc.M) == (1944, 972)
    rng = np.random.default_rng(1944)
    F, q = 64, 0.03
    x = rng.integers(0, 2, (F, oc.N)).astype(np.uint8)
    e = (rng.random((F, oc.N)) < q).astype(np.uint8)
    syn = np.stack([oc.syndrome(x[f]) for f in range(F)])
    mag = np.float32(np.log((1 - q) / q))
    llr = np.where(x ^ e, -mag, mag).astype(np.float32)
    hard, post, iters, ok, _ = oc.batch_flooding_f32(llr, syn, rule=O.RULE_SPA, n_ite=20, early_stop=True)
    assert ok.all() and (hard == x).all()
    assert 2 <= iters.mean() <= 8 and iters.max() <= 20
    assert (np.sign(post) == np.where(x, -1, 1)).all()
