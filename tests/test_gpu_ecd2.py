"""BASELINE config 5 on the real daemon: two processes of the reference's ecd2 (integration/ecd2_ldpc.patch applied, built
here by `make -C oracle ref_ecd2` into oracle/_ref/ecd2_ldpc, which travels to the GPU box) joined by FIFOs, the LDPC
algorithm slot selected; both sides must write the same final key, and the notify line must report the leakage of the
LDPC protocol.  Beside it the unmodified Cascade slot of the same binary, as the reference's own baseline."""
import os
import re
import struct
import subprocess
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "ecd2_ldpc")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/ecd2_ldpc not built")]


def run_pair(tmp, env, n, qber, seed, timeout=120):
    w = tmp / ("run%d" % seed)
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2, n).astype(np.uint8)
    b = a ^ (rng.random(n) < qber)
    for side, bits in (("A", a), ("B", b)):
        for sub in ("raw", "fin"):
            os.makedirs(w / side / sub)
        words = np.packbits(np.concatenate([bits, np.zeros((-n) % 32, np.uint8)])).view(">u4").astype("<u4")
        (w / side / "raw" / "b0b80000").write_bytes(struct.pack("<iIIi", 3, 0xb0b80000, n, 1) + words.tobytes())
        for f in ("cmd", "q"):
            os.mkfifo(w / side / f)
    os.mkfifo(w / "AB")
    os.mkfifo(w / "BA")
    procs = []
    for side, s, r in (("A", "AB", "BA"), ("B", "BA", "AB")):
        d = w / side
        procs.append(subprocess.Popen(["stdbuf", "-o0", EXE, "-c", str(d / "cmd"), "-s", str(w / s), "-r", str(w / r), "-d", str(d / "raw"),
                                       "-f", str(d / "fin"), "-l", str(d / "notify"), "-q", str(d / "resp"), "-Q", str(d / "q"), "-V", "5"],
                                      stdout=open(d / "out", "w"), stderr=subprocess.STDOUT, env=dict(os.environ, **env)))
    time.sleep(0.5)
    t0 = time.time()
    with open(w / "A" / "cmd", "w") as f:
        f.write("0xb0b80000 1\n")
    fin = [w / "A" / "fin" / "b0b80000", w / "B" / "fin" / "b0b80000"]
    while time.time() - t0 < timeout and not all(p.exists() and p.stat().st_size > 0 for p in fin) and all(p.poll() is None for p in procs):
        time.sleep(0.05)
    wall = time.time() - t0
    time.sleep(0.3)
    for p in procs:
        if p.poll() is None:
            p.terminate()
        p.wait(timeout=10)
    notify = (w / "A" / "notify").read_text() if (w / "A" / "notify").exists() else ""
    return fin, notify, wall, (w / "A" / "out").read_text()[-2000:], (w / "B" / "out").read_text()[-2000:]


@pytest.mark.parametrize("alg", ["3", "4"])            # ALG_LDPC_CONTINUE_ROLES / ALG_LDPC_FLIP_ROLES
def test_two_ecd2_daemons_reconcile_with_ldpc(tmp_path, data_dir, alg):
    env = {"ECD2_EC_ALGORITHM": alg, "ECD2_LDPC_BASE_QC": "%s/NR_1_1_384.qc" % data_dir}
    fin, notify, wall, oa, ob = run_pair(tmp_path, env, 40000, 0.03, seed=int(alg))
    assert all(f.exists() and f.stat().st_size > 100 for f in fin), (oa, ob)
    assert fin[0].read_bytes() == fin[1].read_bytes()
    m = re.search(r"initial bit number: (\d+) final bit number: (\d+) error rate: ([0-9.]+) leaked bits in EC: (\d+)", notify)
    assert m, notify
    initial, final, err, leaked = int(m.group(1)), int(m.group(2)), float(m.group(3)), int(m.group(4))
    assert initial == 40000 and 0.02 < err < 0.045 and final > 10000
    assert leaked >= 4 * 384 * 4                                # at least four parity block rows of every frame
    print("LDPC alg %s: final %d bits, leaked %d, error rate %.4f, wall %.2f s" % (alg, final, leaked, err, wall))


def test_cascade_slot_of_the_same_binary(tmp_path):
    fin, notify, wall, oa, ob = run_pair(tmp_path, {}, 40000, 0.03, seed=9)
    assert all(f.exists() and f.stat().st_size > 100 for f in fin), (oa, ob)
    assert fin[0].read_bytes() == fin[1].read_bytes()
    print("Cascade: %s wall %.2f s" % (notify.strip(), wall))
