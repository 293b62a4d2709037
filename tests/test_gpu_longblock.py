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
    assert dec.kernel_name == ("flooding_csr" if rule == "spa" else "flooding_qc")
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
