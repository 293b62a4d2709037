import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def q():
    """the product package (ctypes binding of libqldpc_b200.so)"""
    return importlib.import_module("qcrypto-ldpc_b200")


@pytest.fixture(scope="session")
def O():
    """the CPU oracle (test infrastructure)"""
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def data_dir():
    return os.path.join(ROOT, "qcrypto-ldpc_b200", "data")


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(ROOT, "tests", "golden", "kat_pegreg504x1008.json")) as f:
        return json.load(f)


def make_frames(O, ocode, F, qber, mag, known_mag, mode, seed):
    """Synthetic sifted-key frames: returns int LLRs [F,N], syndrome bits [F,M] or None, truth [F,N].
    mode 'parity'  : the reference's send-parity formulation (info noisy, parity known, zero syndrome)
    mode 'syndrome': all N bits noisy, syndrome of Alice's word given."""
    rng = np.random.default_rng(seed)
    N, M, K = ocode.N, ocode.M, ocode.K
    llr = np.zeros((F, N), dtype=np.int32)
    truth = np.zeros((F, N), dtype=np.uint8)
    syn = None if mode == "parity" else np.zeros((F, M), dtype=np.uint8)
    for f in range(F):
        if mode == "parity":
            msg = rng.integers(0, 2, K).astype(np.uint8)
            cw = ocode.nr_encode(msg)
            e = (rng.random(K) < qber).astype(np.uint8)
            llr[f, :K] = np.where(cw[:K] ^ e, -mag, mag)
            llr[f, K:] = np.where(cw[K:], -known_mag, known_mag)
            truth[f] = cw
        else:
            x = rng.integers(0, 2, N).astype(np.uint8)
            e = (rng.random(N) < qber).astype(np.uint8)
            syn[f] = ocode.syndrome(x)
            llr[f] = np.where(x ^ e, -mag, mag)
            truth[f] = x
    return llr, syn, truth
