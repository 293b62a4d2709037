#!/usr/bin/env python3
"""Golden vectors of the privacy amplification (errorcorrection/subcomponents/priv_amp.c:186-218), generated with the
REFERENCE's own PRNG: oracle/_ref/librnd_ref.so is errorcorrection/subcomponents/rnd.c compiled where it lies
(`make -C oracle ref`, this container only).  Writes tests/golden/pa_golden.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402


def main():
    assert O.ref_rnd() is not None, "build oracle/_ref first: make -C oracle ref"
    rng = np.random.default_rng(20261018)
    cases = []
    for workbits, final_bits in ((1, 1), (31, 7), (32, 32), (33, 20), (1000, 333), (4097, 2500), (40000, 23162)):
        nw = (workbits + 31) // 32
        key = rng.integers(0, 2**32, nw, dtype=np.uint64).astype(np.uint32)
        seed = int(rng.integers(1, 2**32))
        out = O.privacy_amplify(key, workbits, final_bits, seed, use_ref_prng=True)
        cases.append({"workbits": workbits, "final_bits": final_bits, "seed": seed,
                      "key": [int(x) for x in key], "final_key": [int(x) for x in out]})
    # the PRNG itself: 8 successive 32-step words from a fixed seed
    import ctypes as C
    st = C.c_uint32(0xb0b80000)
    R = O.ref_rnd()
    words = [int(R.rnd_getPrngValue2_32(C.byref(st))) for _ in range(8)]
    with open(os.path.join(ROOT, "tests", "golden", "pa_golden.json"), "w") as f:
        json.dump({"source": "errorcorrection/subcomponents/rnd.c (reference, compiled) + priv_amp.c:186-218 loop",
                   "prng_seed": 0xb0b80000, "prng_words": words, "cases": cases}, f)
    print("wrote pa_golden.json:", len(cases), "cases")


if __name__ == "__main__":
    main()
