"""After reconciliation: privacy amplification bit-exact with the reference's priv_amp.c / rnd.c, and
the per-frame confirmation CRC.  Oracle pins: golden vectors generated with the reference's own compiled rnd.c
(tests/golden/make_pa_golden.py) and, where oracle/_ref is present, the reference PRNG itself."""
import ctypes as C
import json
import os
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "pa_golden.json")) as f:
        return json.load(f)


def test_oracle_prng_matches_reference_words(O, golden):
    st = C.c_uint32(golden["prng_seed"])
    got = [int(O.lib().ora_prng32(C.byref(st))) for _ in golden["prng_words"]]
    assert got == golden["prng_words"]


def test_oracle_pa_reproduces_reference_golden_vectors(O, golden):
    for c in golden["cases"]:
        out = O.privacy_amplify(np.array(c["key"], dtype=np.uint32), c["workbits"], c["final_bits"], c["seed"])
        assert [int(x) for x in out] == c["final_key"], (c["workbits"], c["final_bits"])


def test_oracle_pa_against_compiled_reference_prng(O):
    if O.ref_rnd() is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(5)
    for workbits, final_bits in ((77, 50), (2049, 1000), (9999, 4000)):
        key = rng.integers(0, 2**32, (workbits + 31) // 32, dtype=np.uint64).astype(np.uint32)
        seed = int(rng.integers(0, 2**32))
        assert (O.privacy_amplify(key, workbits, final_bits, seed) ==
                O.privacy_amplify(key, workbits, final_bits, seed, use_ref_prng=True)).all()


def test_oracle_crc_is_zlib_crc32(O):
    rng = np.random.default_rng(6)
    for n in (0, 1, 7, 264):
        w = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
        assert O.crc32_words(w) == zlib.crc32(w.astype(">u4").tobytes())


@pytest.mark.gpu
def test_gpu_pa_matches_golden_and_oracle(q, O, golden):
    # the reference-generated vectors, one launch
    cases = golden["cases"]
    W = max(len(c["key"]) for c in cases)
    keys = np.zeros((len(cases), W), np.uint32)
    for b, c in enumerate(cases):
        keys[b, :len(c["key"])] = c["key"]
    out = q.privacy_amplify(keys, [c["workbits"] for c in cases], [c["final_bits"] for c in cases], [c["seed"] for c in cases])
    for b, c in enumerate(cases):
        n = (c["final_bits"] + 31) // 32
        assert [int(x) for x in out[b, :n]] == c["final_key"], (c["workbits"], c["final_bits"])
        assert not out[b, n:].any()
    # random ragged batch incl. empty outputs, garbage beyond workbits in the last key word, ecd2's largest block
    rng = np.random.default_rng(8)
    shapes = [(1, 0), (5, 5), (63, 64 - 1), (64, 1), (65, 33), (4096, 4095), (12345, 6789), (65535, 9000)]
    W = (65535 + 31) // 32
    keys = rng.integers(0, 2**32, (len(shapes), W), dtype=np.uint64).astype(np.uint32)
    seeds = rng.integers(0, 2**32, len(shapes), dtype=np.uint64).astype(np.uint32)
    out = q.privacy_amplify(keys, [s[0] for s in shapes], [s[1] for s in shapes], seeds)
    for b, (wb, fb) in enumerate(shapes):
        want = O.privacy_amplify(keys[b, :(wb + 31) // 32], wb, fb, int(seeds[b]))
        assert (out[b, :(fb + 31) // 32] == want).all(), (wb, fb)


@pytest.mark.gpu
def test_gpu_pa_full_size_block_property(q, O):
    """65 535-bit block -> 40 000 final bits: linear in the key (PA(a) ^ PA(b) == PA(a ^ b)), oracle sampled on a prefix"""
    rng = np.random.default_rng(9)
    W = 2048
    a = rng.integers(0, 2**32, W, dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, 2**32, W, dtype=np.uint64).astype(np.uint32)
    keys = np.stack([a, b, a ^ b])
    out = q.privacy_amplify(keys, [65535] * 3, [40000] * 3, [0xdeadbeef] * 3)
    assert (out[0] ^ out[1] == out[2]).all()
    want = O.privacy_amplify(a, 65535, 640, 0xdeadbeef)            # first 640 final bits on the CPU
    assert (out[0, :20] == want).all()


@pytest.mark.gpu
def test_gpu_crc32_frames(q, O):
    rng = np.random.default_rng(10)
    bits = rng.integers(0, 2**32, (37, 816), dtype=np.uint64).astype(np.uint32)
    got = q.crc32_frames(bits, words_per_frame=264)
    for f in range(37):
        assert int(got[f]) == zlib.crc32(bits[f, :264].astype(">u4").tobytes()) == O.crc32_words(bits[f, :264])
    assert (q.crc32_frames(bits) == [zlib.crc32(bits[f].astype(">u4").tobytes()) for f in range(37)]).all()


def test_postproc_no_cpu_fallback(q):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    keys = np.zeros((1, 4), np.uint32)
    with pytest.raises(q.QldpcError):
        q.privacy_amplify(keys, [100], [50], [1])
    with pytest.raises(q.QldpcError):
        q.crc32_frames(keys)
