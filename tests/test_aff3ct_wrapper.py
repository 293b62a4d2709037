"""The AFF3CT-shaped C++ wrapper (qcrypto-ldpc_b200/host/qldpc_aff3ct.hpp) and the ported driver step."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "qcrypto-ldpc_b200", "host")


@pytest.fixture(scope="module")
def driver(tmp_path_factory, q):
    exe = str(tmp_path_factory.mktemp("drv") / "driver_siho")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", HOST,
                           os.path.join(HOST, "driver_siho.cpp"), "-o", exe, q.LIB_PATH,
                           "-Wl,-rpath," + os.path.dirname(q.LIB_PATH)])
    return exe


def test_driver_builds_and_fails_loudly_without_gpu(driver, data_dir, tmp_path, kat):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    llr = tmp_path / "llr.txt"
    llr.write_text(" ".join("%.2f" % v for v in kat["llrs"]))
    p = subprocess.run([driver, "%s/PEGReg504x1008.alist" % data_dir, str(llr), "504", "504", "10"], capture_output=True, text=True)
    assert p.returncode == 3 and "no sm_100 CUDA device" in p.stderr      # tools::runtime_error, no CPU fallback
    p = subprocess.run([driver, "/nonexistent.alist", str(llr), "504", "504", "10"], capture_output=True, text=True)
    assert p.returncode == 3 and "LDPC_matrix_handler::read" in p.stderr
    assert subprocess.run([driver], capture_output=True).returncode == 2


@pytest.mark.gpu
def test_driver_replays_reference_kat(driver, data_dir, tmp_path, kat):
    """decode_siho(llrs, dec_bits) through the wrapper == the reference's decoded[504] ("main.cpp (alist)":443-462)"""
    llr = tmp_path / "llr.txt"
    frames = [kat["llrs"], [-v for v in kat["llrs"]]]        # two frames back to back (n_frames = 2)
    llr.write_text(" ".join("%.2f" % v for f in frames for v in f))
    p = subprocess.run([driver, "%s/PEGReg504x1008.alist" % data_dir, str(llr), "504", "504", "10"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    lines = p.stdout.strip().split("\n")
    assert len(lines) == 2
    want = "".join(str(b) for b in kat["decoded"])
    assert lines[0].split()[0] == want and "iters=6 ok=1" in lines[0]
    # the negated frame is the complement coset: with zero syndrome it is a different decoding problem, only shape is checked
    assert len(lines[1].split()[0]) == 504


@pytest.mark.gpu
@pytest.mark.parametrize("rule,orule,kw", [("ms", "NMS", dict(norm=1.0)), ("nms:0.75", "NMS", dict(norm=0.75)),
                                            ("oms:0.5", "OMS", dict(offset=0.5))])
def test_driver_rules_match_oracle(driver, O, data_dir, tmp_path, rule, orule, kw):
    oc = O.Code.from_qc("%s/NR_1_1_24.qc" % data_dir)
    rng = np.random.default_rng(12)
    msg = rng.integers(0, 2, oc.K).astype(np.uint8)
    cw = oc.nr_encode(msg)
    e = np.zeros(oc.N, np.uint8)
    e[:oc.K] = rng.random(oc.K) < 0.02
    llrs = np.where(cw ^ e, -3.89, 3.89).astype(np.float32)
    llrs[oc.K:] = np.where(cw[oc.K:], -23.02585, 23.02585)              # confirmed parity (BOOT/src/main.cpp:351-354)
    f = tmp_path / "llr.txt"
    f.write_text(" ".join(repr(float(v)) for v in llrs))
    for sched, fn in (([], oc.decode_flooding_f32), ([rule, "layered"], oc.decode_layered_f32)):
        args = [driver, "%s/NR_1_1_24.qc" % data_dir, str(f), "0", str(oc.K), "10"] + (sched if sched else [rule])
        p = subprocess.run(args, capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        hard, _, it, ok = fn(llrs, None, rule=getattr(O, "RULE_" + orule), n_ite=10, **kw)
        bits, its, okk = p.stdout.split()
        assert bits == "".join(str(b) for b in hard[:oc.K]) and its == "iters=%d" % it and okk == "ok=%d" % ok


@pytest.mark.gpu
def test_decode_siho_reaches_the_int8_kernel(driver, O, data_dir, tmp_path):
    """Decoder_LDPC_BP<int, int8_t> (the Q template parameter of Decoder_SISO_SIHO<B,Q>, BOOT/src/main.cpp:113)
    on BG1 Z=384, layered NMS 6/8 -> the streamed int8 kernel of the headline benchmark; bits, iteration counts and flags equal
    to the oracle's.  A factor that is not k/8 and SPA with an integer Q raise tools::invalid_argument."""
    oc = O.Code.from_qc("%s/NR_1_1_384.qc" % data_dir)
    rng = np.random.default_rng(21)
    F = 6
    llr = np.zeros((F, oc.N), np.int8)
    for f in range(F):
        cw = oc.nr_encode(rng.integers(0, 2, oc.K).astype(np.uint8))
        e = np.zeros(oc.N, np.uint8)
        e[:oc.K] = rng.random(oc.K) < 0.035
        llr[f] = np.where(cw ^ e, -13, 13)
        llr[f, oc.K:] = np.where(cw[oc.K:], -31, 31)
    fn = tmp_path / "llr.txt"
    fn.write_text(" ".join(str(int(v)) for v in llr.ravel()))
    p = subprocess.run([driver, "%s/NR_1_1_384.qc" % data_dir, str(fn), "0", str(oc.K), "10", "nms:0.75", "layered", "i8"],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "kernel=layered_i8s_zpack4" in p.stderr
    hard, oit, ook, _ = oc.batch_layered_fixed_i8(llr, None, rule=O.RULE_NMS, n_ite=10, early_stop=True, norm_eighths=6)
    lines = p.stdout.strip().split("\n")
    assert len(lines) == F
    for f in range(F):
        bits, its, okk = lines[f].split()
        assert bits == "".join(str(b) for b in hard[f, :oc.K]) and its == "iters=%d" % oit[f] and okk == "ok=%d" % ook[f]
    # int16 Q on the flooding schedule, and the argument checks of the integer rules
    p = subprocess.run([driver, "%s/NR_1_1_384.qc" % data_dir, str(fn), "0", str(oc.K), "10", "oms:1", "flooding", "i16"],
                       capture_output=True, text=True)
    assert p.returncode == 0 and "kernel=flooding" in p.stderr, p.stderr
    h16, _, it16, ok16 = oc.decode_flooding_fixed(llr[0].astype(np.int32), None, rule=O.RULE_OMS, n_ite=10, early_stop=True, offset=1,
                                                 vmax=32767)
    bits, its, okk = p.stdout.strip().split("\n")[0].split()
    assert bits == "".join(str(b) for b in h16[:oc.K]) and its == "iters=%d" % it16 and okk == "ok=%d" % ok16
    for bad in ("nms:0.8", "spa"):
        p = subprocess.run([driver, "%s/NR_1_1_384.qc" % data_dir, str(fn), "0", str(oc.K), "10", bad, "layered", "i8"],
                           capture_output=True, text=True)
        assert p.returncode == 4 and "error:" in p.stderr, (bad, p.returncode, p.stderr)


@pytest.fixture(scope="module")
def driver_5gqc(tmp_path_factory, q):
    exe = str(tmp_path_factory.mktemp("drv5g") / "driver_5gqc")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", HOST,
                           os.path.join(HOST, "driver_5gqc.cpp"), "-o", exe, q.LIB_PATH,
                           "-Wl,-rpath," + os.path.dirname(q.LIB_PATH)])
    return exe


def test_5gqc_driver_builds_and_fails_loudly_without_gpu(driver_5gqc, data_dir):
    import torch
    assert subprocess.run([driver_5gqc], capture_output=True).returncode == 2
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    p = subprocess.run([driver_5gqc, "%s/NR_1_0_2.qc" % data_dir, "1", "1.0"], capture_output=True, text=True)
    assert p.returncode == 3 and "no sm_100 CUDA device" in p.stderr


@pytest.mark.gpu
def test_5gqc_driver_reproduces_the_recorded_sweep(driver_5gqc, data_dir):
    """the C++ port of the (5g-qc) driver loop (block puncturing to a target efficiency, QBER loop 0 .. 0.11, flooding SPA,
    n_ite 10) with the command line that produced README_LDPC.md:937-974 (NR_1_0_2.qc, expansion_factor 1, efficiency 1.0):
    the recorded puncture counts are reproduced exactly, the frame-error counts statistically (2 000 frames per step)"""
    import re
    import refpins
    rows = refpins.recorded()["readme_ldpc_nr_1_0_2"]["rows"]
    p = subprocess.run([driver_5gqc, "%s/NR_1_0_2.qc" % data_dir, "1", "1.0", "2000"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    punct = [int(x) for x in re.findall(r"Punct\. Bits \(Round down\): (\d+)\|", p.stdout)]
    res = re.findall(r"^\s+([0-9.]+) \|\|\s+(\d+) \|\s+(\d+) \|\s+(\d+) \|", p.stdout, re.M)
    assert len(punct) == len(res) == 12
    assert punct == [r["punctured_bits"] for r in rows]
    inside = 0
    for (qber, frames, be, fe), row in list(zip(res, rows))[1:]:
        assert abs(float(qber) - row["qber"]) < 1e-3 and int(frames) == 2000
        assert refpins.consistent(int(fe), 2000, row["frame_errors"], row["frames"]), (qber, fe, row)
        lo, hi = refpins.clopper_pearson(row["frame_errors"], row["frames"])
        inside += lo <= int(fe) / 2000 <= hi
    assert inside >= 9


@pytest.fixture(scope="module")
def encoders_exe(tmp_path_factory, q):
    exe = str(tmp_path_factory.mktemp("enc") / "test_encoders")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", HOST,
                           os.path.join(HOST, "test_encoders.cpp"), "-o", exe, q.LIB_PATH,
                           "-Wl,-rpath," + os.path.dirname(q.LIB_PATH)])
    return exe


def test_encoder_classes_build_and_fail_loudly_without_gpu(encoders_exe, data_dir, tmp_path, kat):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    (tmp_path / "d.txt").write_text(" ".join(str(b) for b in kat["data"]))
    (tmp_path / "e.txt").write_text(" ".join(str(b) for b in kat["encoded"]))
    p = subprocess.run([encoders_exe, "%s/PEGReg504x1008.alist" % data_dir, "%s/G_PEGReg504x1008.alist" % data_dir,
                        str(tmp_path / "d.txt"), str(tmp_path / "e.txt")], capture_output=True, text=True)
    assert p.returncode == 3 and "no sm_100 CUDA device" in p.stderr


@pytest.mark.gpu
def test_encoder_classes_replay_kat_e(encoders_exe, data_dir, tmp_path, kat):
    """Encoder_LDPC<B>(K, N, G) and Encoder_LDPC_from_H<B>(K, N, H) of the AFF3CT face reproduce the reference's encoded[1008]
    ("main.cpp (alist)":443-462) and are systematic at get_info_bits_pos()"""
    (tmp_path / "d.txt").write_text(" ".join(str(b) for b in kat["data"]))
    (tmp_path / "e.txt").write_text(" ".join(str(b) for b in kat["encoded"]))
    p = subprocess.run([encoders_exe, "%s/PEGReg504x1008.alist" % data_dir, "%s/G_PEGReg504x1008.alist" % data_dir,
                        str(tmp_path / "d.txt"), str(tmp_path / "e.txt")], capture_output=True, text=True)
    assert p.returncode == 0 and "mismatches 0" in p.stdout, (p.stdout, p.stderr)
