"""integration/ecd2_ldpc.patch against the reference's ecd2 daemon: the patch applies to a scratch copy of
/root/reference/errorcorrection, `make` builds and links ecd2 against libqldpc_b200.so, the unmodified Cascade path still
reconciles a block, and with the LDPC slot selected a daemon without an sm_100 device reports the reference's own error
81 instead of crashing.  Skipped where /root/reference is absent (the GPU box); on a GPU the same two daemons reconcile the
block with LDPC (tests/test_ecd2_patch.py::test_ldpc_two_daemons_on_gpu needs the reference tree and is skipped there too)."""
import os
import shutil
import struct
import subprocess
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF + "/errorcorrection"), reason="reference tree not present")


@pytest.fixture(scope="module")
def ecd2_tree(tmp_path_factory, q):
    d = tmp_path_factory.mktemp("ecd2")
    shutil.copytree(REF + "/errorcorrection", d / "errorcorrection", ignore=shutil.ignore_patterns("ldpc_examples", "readme_imgs"))
    shutil.copytree(REF + "/packetheaders", d / "packetheaders")
    subprocess.check_call(["patch", "-p1", "--binary", "-i", os.path.join(ROOT, "integration", "ecd2_ldpc.patch")], cwd=d)
    p = subprocess.run(["make", "QLDPC_ROOT=" + ROOT], cwd=d / "errorcorrection", capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    assert " error:" not in p.stderr and "ldpc_reconcile.c" not in p.stderr     # our file compiles without a warning
    return d


def test_patch_is_what_the_generator_emits(tmp_path):
    """the committed patch is reproducible from integration/make_ecd2_patch.py + integration/ecd2/*"""
    committed = open(os.path.join(ROOT, "integration", "ecd2_ldpc.patch"), "rb").read()
    subprocess.check_call(["python", os.path.join(ROOT, "integration", "make_ecd2_patch.py")], stdout=subprocess.DEVNULL)
    assert open(os.path.join(ROOT, "integration", "ecd2_ldpc.patch"), "rb").read() == committed


def test_ecd2_links_against_the_library(ecd2_tree):
    exe = ecd2_tree / "errorcorrection" / "ecd2"
    assert exe.exists()
    ldd = subprocess.run(["ldd", str(exe)], capture_output=True, text=True).stdout
    assert "libqldpc_b200.so" in ldd and "not found" not in ldd
    syms = subprocess.run(["nm", "-u", str(exe)], capture_output=True, text=True).stdout
    for s in ("qldpc_ecd2_open", "qldpc_ecd2_initiate", "qldpc_ecd2_handle", "qldpc_ecd2_packet_data"):
        assert s in syms


def _two_party_run(tree, env, n=20000, qber=0.03, timeout=60):
    """two daemons joined by FIFOs, one stream-3 block each, `epoch 1` on Alice's command pipe"""
    w = tree / ("run%d" % int(time.time() * 1e3))
    rng = np.random.default_rng(3)
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
    exe = str(tree / "errorcorrection" / "ecd2")
    procs = []
    for side, s, r in (("A", "AB", "BA"), ("B", "BA", "AB")):
        d = w / side
        procs.append(subprocess.Popen(["stdbuf", "-o0", exe, "-c", str(d / "cmd"), "-s", str(w / s), "-r", str(w / r), "-d", str(d / "raw"), "-f", str(d / "fin"),
                                       "-l", str(d / "notify"), "-q", str(d / "resp"), "-Q", str(d / "q"), "-V", "5"],
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=dict(os.environ, **env)))
    time.sleep(0.5)
    with open(w / "A" / "cmd", "w") as f:
        f.write("0xb0b80000 1\n")
    t0 = time.time()
    fin = [w / "A" / "fin" / "b0b80000", w / "B" / "fin" / "b0b80000"]
    while time.time() - t0 < timeout and not all(p.exists() and p.stat().st_size > 0 for p in fin) and all(p.poll() is None for p in procs):
        time.sleep(0.2)
    time.sleep(0.3)
    out = []
    for p in procs:
        if p.poll() is None:
            p.terminate()
        try:
            out.append(p.communicate(timeout=5)[0])
        except subprocess.TimeoutExpired:
            p.kill()
            out.append(p.communicate()[0])
    return w, fin, out, [p.returncode for p in procs]


def test_cascade_path_is_untouched(ecd2_tree):
    """default algorithm choice (Cascade/BICONF): the patched daemon still produces identical final keys on both sides"""
    w, fin, out, rc = _two_party_run(ecd2_tree, {})
    assert all(f.exists() for f in fin), out
    assert fin[0].read_bytes() == fin[1].read_bytes() and fin[0].stat().st_size > 100


def test_ldpc_slot_without_gpu_reports_error_81(ecd2_tree, data_dir):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    env = {"ECD2_EC_ALGORITHM": "3", "ECD2_LDPC_BASE_QC": "%s/NR_1_1_384.qc" % data_dir}
    w, fin, out, rc = _two_party_run(ecd2_tree, env, timeout=15)
    joined = "\n".join(out)
    assert "LDPC reconciliation unavailable" in joined and "Unsupported functionality" in joined
    assert not any(f.exists() and f.stat().st_size > 0 for f in fin)          # no key without a decoder, and no crash
    assert all(r not in (-11, -6, -7, -8) for r in rc)                           # neither daemon crashed (SIGSEGV / SIGABRT / ...)
    assert (-81) % 256 in rc                                                     # the initiator left through `return -errorCode`
