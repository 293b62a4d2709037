"""The C-ABI library without a GPU: it loads, exports every symbol include/qldpc.h declares, parses the
reference's matrix formats exactly like the oracle, validates arguments, and refuses to compute without
a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "qldpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qldpc_[a-z0-9_]+)\s*\(", src)))


def test_ecd2_face_exports_every_declared_symbol(q):
    """include/qldpc_ecd2.h: the C face the patched ecd2 daemon binds (integration/ecd2_ldpc.patch); plain C"""
    import subprocess
    import tempfile
    src = open(os.path.join(ROOT, "include", "qldpc_ecd2.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = sorted(set(re.findall(r"\b(qldpc_ecd2_[a-z0-9_]+)\s*\(", src)))
    assert len(declared) == 8
    for name in declared:
        assert hasattr(q.lib(), name), "libqldpc_b200.so does not export " + name
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write('#include "qldpc_ecd2.h"\nint main(void){qldpc_ecd2_config c; qldpc_ecd2 *x = 0; qldpc_ecd2_config_default(&c);'
                           ' return qldpc_ecd2_open(&c, &x) == 81 ? 0 : 1;}\n')     # no base graph -> 81, no crash
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), c, "-o", exe,
                               q.LIB_PATH, "-Wl,-rpath," + os.path.dirname(q.LIB_PATH)])
        assert subprocess.call([exe]) == 0


def test_library_exports_every_declared_symbol(q):
    L = q.lib()
    declared = _header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), "libqldpc_b200.so does not export " + name
    assert sorted(q.ABI_SYMBOLS) == declared
    assert L.qldpc_version() == 200


def test_signatures_are_plain_c(q):
    """no torch / C++ types cross the boundary: the header compiles as C."""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write('#include "qldpc.h"\nint main(void){qldpc_decoder_config c; qldpc_decoder_config_default(&c); return c.max_iter==100?0:1;}\n')
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                               q.LIB_PATH, "-Wl,-rpath," + os.path.dirname(q.LIB_PATH)])
        assert subprocess.call([exe]) == 0


@pytest.mark.parametrize("name", ["PEGReg504x1008.alist", "20.alist", "1998.5.3.2665.alist"])
def test_alist_parser_matches_oracle(q, O, data_dir, name):
    code = q.Code.from_alist("%s/%s" % (data_dir, name))
    oc = O.Code.from_alist("%s/%s" % (data_dir, name))
    H = oc.dense()
    assert (code.n, code.m, code.edges, code.z) == (oc.N, oc.M, oc.E, 0)
    assert code.k == oc.N - oc.M
    assert code.max_chk_degree == H.sum(axis=1).max() and code.max_var_degree == H.sum(axis=0).max()


@pytest.mark.parametrize("name", ["NR_1_1_384.qc", "NR_1_1_192.qc", "NR_2_3_112.qc", "test.qc", "test2.qc", "NR_1_0_2.qc",
                                  "wifi_n1944_r12.qc"])
def test_qc_parser_matches_oracle(q, O, data_dir, name):
    code = q.Code.from_qc_file("%s/%s" % (data_dir, name))
    oc = O.Code.from_qc("%s/%s" % (data_dir, name))
    assert (code.n, code.m, code.edges, code.z, code.base_rows, code.base_cols) == (oc.N, oc.M, oc.E, oc.Z, oc.brows, oc.bcols)
    b = oc.base
    c2 = q.Code.from_qc(b, oc.Z)
    assert (c2.n, c2.m, c2.edges) == (code.n, code.m, code.edges)
    assert code.max_chk_degree == (b >= 0).sum(axis=1).max() and code.max_var_degree == (b >= 0).sum(axis=0).max()


def test_csr_constructor_and_info_bits(q, O, data_dir):
    oc = O.Code.from_alist("%s/PEGReg504x1008.alist" % data_dir)
    code = q.Code.from_csr(oc.N, oc.M, oc.row_ptr, oc.col_idx)
    assert (code.n, code.m, code.k, code.edges) == (1008, 504, 504, 3024)
    code.set_info_bits_pos(np.arange(100, 400))
    assert code.k == 300
    with pytest.raises(q.QldpcError):
        code.set_info_bits_pos(np.array([5000]))


def test_error_codes(q, data_dir, tmp_path):
    L = q.lib()
    h = C.c_void_p()
    assert L.qldpc_code_from_alist_file(b"/nonexistent/file.alist", C.byref(h)) == 2       # QLDPC_ERR_IO
    bad = tmp_path / "bad.alist"
    bad.write_text("4 2\n2 2\n1 1 1 1\n2 2\n1 0\n1 0\n2 0\n")                             # truncated
    assert L.qldpc_code_from_alist_file(str(bad).encode(), C.byref(h)) == 3               # QLDPC_ERR_FORMAT
    badqc = tmp_path / "bad.qc"
    badqc.write_text("3 2 5\n\n1 2 3\n1 2\n")
    assert L.qldpc_code_from_qc_file(str(badqc).encode(), C.byref(h)) == 3
    assert L.qldpc_code_from_qc_file(None, C.byref(h)) == 1                               # QLDPC_ERR_ARG
    assert L.qldpc_decode(None, None, None, 0, None, None, None, None) == 1
    assert L.qldpc_strerror(7) == b"no sm_100 CUDA device"
    # integer tiers: SPA is float only; NMS factor must be k/8
    code = q.Code.from_qc_file("%s/NR_1_1_24.qc" % data_dir)
    cfg = q.DecoderConfig()
    L.qldpc_decoder_config_default(C.byref(cfg))
    assert (cfg.schedule, cfg.rule, cfg.dtype, cfg.max_iter, cfg.early_stop, cfg.syndrome_depth) == (0, 0, 0, 100, 1, 1)
    cfg.dtype = q.DTYPE_I8
    d = C.c_void_p()
    assert L.qldpc_decoder_create(code.h, C.byref(cfg), C.byref(d)) == 6                  # QLDPC_ERR_UNSUPPORTED
    al = q.Code.from_alist("%s/20.alist" % data_dir)
    cfg.rule, cfg.schedule = q.RULE_NMS, q.SCHED_LAYERED
    assert L.qldpc_decoder_create(al.h, C.byref(cfg), C.byref(d)) == 6                    # integer layered decoding needs a QC code (float runs on layered_csr)


def test_no_cpu_fallback(q, data_dir):
    """without a CUDA device every compute path fails loudly (QLDPC_ERR_NO_DEVICE), it never computes on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    code = q.Code.from_qc_file("%s/NR_1_1_24.qc" % data_dir)
    with pytest.raises(q.QldpcError) as ei:
        q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, norm_factor=0.75)
    assert ei.value.code == 7


def test_bit_packing_convention(q, O):
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 2, (3, 77)).astype(np.uint8)
    w = q.pack_bits(bits)
    assert w.shape == (3, 3) and (w == O.pack_bits_msb(bits)).all()
    assert (q.unpack_bits(w, 77) == bits).all()
    one = np.zeros(64, np.uint8)
    one[0] = 1
    assert q.pack_bits(one).tolist() == [0x80000000, 0]      # helpers.h:68
