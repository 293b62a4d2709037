"""Workloads of the reference's own recorded LDPC results (tests/golden/recorded_fer.json), shared by the CPU-oracle and the
GPU pin tests.

(a) ML/BPSK_nrldpc_sim_RM_FP.m:1-37 -- rate-matched NR code (first mbRM block rows / nbRM block columns of the base graph),
    BPSK over AWGN, first 2z LLRs zeroed, `floor(r / rmax * maxqr)` quantiser clipped to [-32, 31], integer layered offset
    min-sum (offset 2, 20 iterations, no early stop), frame error = any of the k message bits wrong (:100-106).
    NOISE LEVEL.  The script as committed computes sigma from the rate k / (n - 2z) (:24).  The table recorded in
    ML/sim_results.m is NOT reproduced by that: the decoder (oracle, a literal transcription of the .m file, and the GPU alike)
    is 0.19 dB (BG1) / 0.41 dB (BG2) better.  Those two offsets are exactly 10 log10((k / (n - 2z)) / (k / n)) for the two
    codes, and with sigma = sqrt(1 / (2 (k / n) EbNo)) -- the punctured 2z columns counted as transmitted -- all eight
    recorded points are reproduced inside their 95 % confidence intervals.  The recorded runs were evidently made with that
    rate definition; `rate_def` selects it ("k_over_n", default for the pin) or the committed script's ("punctured").
(b) "main.cpp (5g-qc)":450-457,497-537 -- NR_1_0_2.qc, flooding SPA, n_ite 10, early stop, BSC; parity bits sent as known
    (+-CONFIRMED_BIT_LLR), the last `bits_to_puncture` bits erased; desired efficiency 1.0, expansion_factor argument 1
    (the recorded puncture counts 88, 85, 83, ... are only reproduced by that argument pair).
"""
import json
import os

import numpy as np
from scipy.stats import beta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIRMED_BIT_LLR = float(-np.log(1e-10 / (1 - 1e-10)))   # "main.cpp (5g-qc)":33


def recorded():
    with open(os.path.join(ROOT, "tests", "golden", "recorded_fer.json")) as f:
        return json.load(f)


def clopper_pearson(k, n, alpha=0.05):
    lo = 0.0 if k == 0 else float(beta.ppf(alpha / 2, k, n - k + 1))
    hi = 1.0 if k == n else float(beta.ppf(1 - alpha / 2, k + 1, n - k))
    return lo, hi


def consistent(k1, n1, k2, n2, alpha=0.05):
    """two binomial samples agree: their Clopper-Pearson (1 - alpha) intervals overlap"""
    a, b = clopper_pearson(k1, n1, alpha), clopper_pearson(k2, n2, alpha)
    return a[0] <= b[1] and b[0] <= a[1]


def rm_fp_geometry(base):
    """BPSK_nrldpc_sim_RM_FP.m:10-22: (kb, nbRM, mbRM) of the rate-1/2 rate-matched code"""
    mb, nb = base.shape
    kb = nb - mb
    nbRM = int(np.ceil(kb / 0.5)) + 2
    return kb, nbRM, nbRM - kb


def rm_fp_frames(base, z, ebno_db, n_frames, encode, seed, rate_def="k_over_n", rmax=3, maxqr=31):
    """-> (msgs [F,k] uint8, llr [F,n] int8).  encode(msgs[F,k]) -> codewords [F, nb*z] of the FULL code (:30-31)."""
    kb, nbRM, mbRM = rm_fp_geometry(base)
    k, n = kb * z, nbRM * z
    rate = k / n if rate_def == "k_over_n" else k / (n - 2 * z)
    sigma = np.sqrt(1.0 / (2.0 * rate * 10 ** (ebno_db / 10.0)))          # :23-24
    rng = np.random.default_rng(seed)
    msgs = rng.integers(0, 2, (n_frames, k)).astype(np.uint8)
    cw = encode(msgs)[:, :n]
    r = (1.0 - 2.0 * cw) + sigma * rng.standard_normal((n_frames, n))     # :33-34
    r[:, :2 * z] = 0                                                       # :36
    rq = np.clip(np.floor(r / rmax * maxqr), -(maxqr + 1), maxqr)          # :38-40
    return msgs, rq.astype(np.int8)


def binary_entropy(q):
    return float(-q * np.log2(q) - (1 - q) * np.log2(1 - q))


def qc5g_puncture(n, k, qber, efficiency=1.0, expansion_factor=1):
    """"main.cpp (5g-qc)":450-452"""
    bits = int(n - k - efficiency * binary_entropy(qber) * k)
    return bits // expansion_factor * expansion_factor


def qc5g_frames(n, k, qber, n_frames, encode, seed, efficiency=1.0):
    """-> (msgs [F,k], llr [F,n] float32, punctured bits): "main.cpp (5g-qc)":497-530 for one QBER step"""
    punct = qc5g_puncture(n, k, qber, efficiency)
    rng = np.random.default_rng(seed)
    msgs = rng.integers(0, 2, (n_frames, k)).astype(np.uint8)
    cw = encode(msgs)
    noisy = cw ^ (rng.random((n_frames, n)) < qber)
    mag = np.float32(np.log((1 - qber) / qber))                            # Modem_OOK_BSC::demodulate
    llr = np.where(noisy, -mag, mag).astype(np.float32)
    llr[:, k:n - punct] = np.where(cw[:, k:n - punct], -CONFIRMED_BIT_LLR, CONFIRMED_BIT_LLR)   # :514-523
    llr[:, n - punct:] = 0                                                 # :527-530
    return msgs, llr, punct
