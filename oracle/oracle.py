"""ctypes binding of the CPU oracle (oracle/qldpc_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqldpc_oracle.so")

RULE_SPA, RULE_NMS, RULE_OMS = 0, 1, 2


def build(force=False):
    src = os.path.join(_HERE, "qldpc_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.ora_code_from_alist_file.restype = vp
        L.ora_code_from_alist_file.argtypes = [C.c_char_p]
        L.ora_code_from_qc_file.restype = vp
        L.ora_code_from_qc_file.argtypes = [C.c_char_p]
        L.ora_code_from_nr_txt.restype = vp
        L.ora_code_from_nr_txt.argtypes = [C.c_char_p, ci]
        L.ora_code_from_base.restype = vp
        L.ora_code_from_base.argtypes = [vp, ci, ci, ci]
        L.ora_code_from_csr.restype = vp
        L.ora_code_from_csr.argtypes = [ci, ci, vp, vp]
        L.ora_code_free.argtypes = [vp]
        for n in ("N", "M", "E", "Z", "brows", "bcols"):
            f = getattr(L, "ora_code_" + n)
            f.restype, f.argtypes = ci, [vp]
        for n in ("row_ptr", "col_idx", "base"):
            f = getattr(L, "ora_code_" + n)
            f.restype, f.argtypes = C.POINTER(ci), [vp]
        L.ora_syndrome.argtypes = [vp, vp, vp]
        L.ora_nr_encode.restype = ci
        L.ora_nr_encode.argtypes = [vp, vp, vp]
        for n in ("ora_decode_flooding_f32", "ora_decode_layered_f32"):
            f = getattr(L, n)
            f.restype = ci
            f.argtypes = [vp, vp, vp, ci, ci, ci, ci, cf, cf, vp, vp, vp]
        L.ora_decode_layered_fixed.restype = ci
        L.ora_decode_layered_fixed.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp]
        L.ora_decode_flooding_fixed.restype = ci
        L.ora_decode_flooding_fixed.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, vp, vp, vp]
        L.ora_batch_layered_fixed_i8.restype = ci
        L.ora_batch_layered_fixed_i8.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp, ci]
        L.ora_batch_flooding_f32.restype = ci
        L.ora_batch_flooding_f32.argtypes = [vp, vp, vp, ci, ci, ci, ci, cf, cf, vp, vp, vp, vp, ci]
        L.ora_normalize_eighths.restype = ci
        L.ora_normalize_eighths.argtypes = [ci, ci]
        L.ora_privacy_amplify.restype = None
        L.ora_privacy_amplify.argtypes = [vp, ci, ci, C.c_uint32, vp, vp]
        L.ora_prng32.restype = C.c_uint32
        L.ora_prng32.argtypes = [vp]
        L.ora_crc32_words.restype = C.c_uint32
        L.ora_crc32_words.argtypes = [vp, ci]
        _lib = L
    return _lib


_REF_RND = os.path.join(_HERE, "_ref", "librnd_ref.so")


def ref_rnd():
    """the reference's own PRNG (errorcorrection/subcomponents/rnd.c compiled by `make -C oracle ref`), or None when
    oracle/_ref has not been built (it is built in the container that has /root/reference and travels to the GPU box)"""
    if not os.path.exists(_REF_RND):
        return None
    R = C.CDLL(_REF_RND)
    R.rnd_getPrngValue2_32.restype = C.c_uint32
    R.rnd_getPrngValue2_32.argtypes = [C.c_void_p]
    return R


def privacy_amplify(key_words, workbits, final_bits, seed, use_ref_prng=False):
    """priv_amp.c:186-218 on one block; use_ref_prng=True draws the PRNG words from the reference's rnd.c (oracle/_ref)"""
    key = np.ascontiguousarray(key_words, dtype=np.uint32)
    out = np.zeros(max(1, (final_bits + 31) // 32), dtype=np.uint32)
    fn = None
    if use_ref_prng:
        R = ref_rnd()
        assert R is not None, "oracle/_ref/librnd_ref.so is missing (make -C oracle ref)"
        fn = C.cast(R.rnd_getPrngValue2_32, C.c_void_p)
    lib().ora_privacy_amplify(_p(key), int(workbits), int(final_bits), int(seed) & 0xffffffff, _p(out), fn)
    return out[:(final_bits + 31) // 32]


def crc32_words(words):
    w = np.ascontiguousarray(words, dtype=np.uint32)
    return int(lib().ora_crc32_words(_p(w), int(w.size)))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Code:
    """A parity-check matrix as the oracle sees it (CSR by check, ascending variables)."""

    def __init__(self, handle):
        if not handle:
            raise ValueError("oracle: could not build code")
        self.h = handle
        L = lib()
        self.N, self.M, self.E = L.ora_code_N(handle), L.ora_code_M(handle), L.ora_code_E(handle)
        self.Z, self.brows, self.bcols = L.ora_code_Z(handle), L.ora_code_brows(handle), L.ora_code_bcols(handle)
        self.K = self.N - self.M

    @classmethod
    def from_alist(cls, path):
        return cls(lib().ora_code_from_alist_file(path.encode()))

    @classmethod
    def from_qc(cls, path):
        return cls(lib().ora_code_from_qc_file(path.encode()))

    @classmethod
    def from_nr_txt(cls, path, Z):
        return cls(lib().ora_code_from_nr_txt(path.encode(), Z))

    @classmethod
    def from_base(cls, base, Z):
        base = np.ascontiguousarray(base, dtype=np.int32)
        return cls(lib().ora_code_from_base(_p(base), base.shape[0], base.shape[1], Z))

    @classmethod
    def from_csr(cls, N, M, row_ptr, col_idx):
        rp = np.ascontiguousarray(row_ptr, dtype=np.int32)
        ci = np.ascontiguousarray(col_idx, dtype=np.int32)
        return cls(lib().ora_code_from_csr(N, M, _p(rp), _p(ci)))

    def __del__(self):
        try:
            lib().ora_code_free(self.h)
        except Exception:
            pass

    @property
    def row_ptr(self):
        return np.ctypeslib.as_array(lib().ora_code_row_ptr(self.h), shape=(self.M + 1,)).copy()

    @property
    def col_idx(self):
        return np.ctypeslib.as_array(lib().ora_code_col_idx(self.h), shape=(self.E,)).copy()

    @property
    def base(self):
        if not self.Z:
            return None
        return np.ctypeslib.as_array(lib().ora_code_base(self.h), shape=(self.brows, self.bcols)).copy()

    def dense(self):
        H = np.zeros((self.M, self.N), dtype=np.uint8)
        rp, ci = self.row_ptr, self.col_idx
        for m in range(self.M):
            H[m, ci[rp[m]:rp[m + 1]]] = 1
        return H

    # ---- GF(2) helpers
    def syndrome(self, bits):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        syn = np.zeros(self.M, dtype=np.uint8)
        lib().ora_syndrome(self.h, _p(bits), _p(syn))
        return syn

    def nr_encode(self, msg):
        msg = np.ascontiguousarray(msg, dtype=np.uint8)
        cw = np.zeros(self.N, dtype=np.uint8)
        if lib().ora_nr_encode(self.h, _p(msg), _p(cw)) != 0:
            raise ValueError("nr_encode: not an NR-shaped QC code")
        return cw

    # ---- decoders (single frame)
    def _f32(self, fn, llr, syn, rule, n_ite, early_stop, depth, norm, offset):
        llr = np.ascontiguousarray(llr, dtype=np.float32)
        syn = None if syn is None else np.ascontiguousarray(syn, dtype=np.uint8)
        post = np.zeros(self.N, dtype=np.float32)
        hard = np.zeros(self.N, dtype=np.uint8)
        it = C.c_int(0)
        ok = fn(self.h, _p(llr), _p(syn), rule, n_ite, int(early_stop), depth, norm, offset,
                _p(post), _p(hard), C.addressof(it))
        return hard, post, it.value, ok == 1

    def decode_flooding_f32(self, llr, syn=None, rule=RULE_SPA, n_ite=20, early_stop=True, depth=1,
                            norm=1.0, offset=0.0):
        return self._f32(lib().ora_decode_flooding_f32, llr, syn, rule, n_ite, early_stop, depth, norm, offset)

    def decode_layered_f32(self, llr, syn=None, rule=RULE_NMS, n_ite=20, early_stop=True, depth=1,
                           norm=1.0, offset=0.0):
        return self._f32(lib().ora_decode_layered_f32, llr, syn, rule, n_ite, early_stop, depth, norm, offset)

    def decode_layered_fixed(self, llr, syn=None, rule=RULE_OMS, n_ite=20, early_stop=False, offset=2,
                             norm_eighths=6, msg_max=31, app_max=127):
        llr = np.ascontiguousarray(llr, dtype=np.int32)
        syn = None if syn is None else np.ascontiguousarray(syn, dtype=np.uint8)
        app = np.zeros(self.N, dtype=np.int32)
        hard = np.zeros(self.N, dtype=np.uint8)
        it = C.c_int(0)
        ok = lib().ora_decode_layered_fixed(self.h, _p(llr), _p(syn), rule, n_ite, int(early_stop), offset,
                                            norm_eighths, msg_max, app_max, _p(app), _p(hard), C.addressof(it))
        if ok < 0:
            raise ValueError("layered fixed-point oracle needs a QC code")
        return hard, app, it.value, ok == 1

    def decode_flooding_fixed(self, llr, syn=None, rule=RULE_NMS, n_ite=20, early_stop=True, offset=0,
                              norm_eighths=8, vmax=127):
        llr = np.ascontiguousarray(llr, dtype=np.int32)
        syn = None if syn is None else np.ascontiguousarray(syn, dtype=np.uint8)
        post = np.zeros(self.N, dtype=np.int32)
        hard = np.zeros(self.N, dtype=np.uint8)
        it = C.c_int(0)
        ok = lib().ora_decode_flooding_fixed(self.h, _p(llr), _p(syn), rule, n_ite, int(early_stop), offset,
                                             norm_eighths, vmax, _p(post), _p(hard), C.addressof(it))
        if ok < 0:
            raise ValueError("fixed-point flooding has no SPA rule")
        return hard, post, it.value, ok == 1

    # ---- batched (pthreads over frames)
    def batch_layered_fixed_i8(self, llr, syn=None, rule=RULE_NMS, n_ite=10, early_stop=True, offset=2,
                               norm_eighths=6, msg_max=31, app_max=127, n_threads=0):
        llr = np.ascontiguousarray(llr, dtype=np.int8)
        F = llr.shape[0]
        syn = None if syn is None else np.ascontiguousarray(syn, dtype=np.uint8)
        hard = np.zeros((F, self.N), dtype=np.uint8)
        iters = np.zeros(F, dtype=np.int32)
        ok = np.zeros(F, dtype=np.uint8)
        nt = lib().ora_batch_layered_fixed_i8(self.h, _p(llr), _p(syn), F, rule, n_ite, int(early_stop), offset,
                                              norm_eighths, msg_max, app_max, _p(hard), _p(iters), _p(ok), n_threads)
        return hard, iters, ok.astype(bool), nt

    def batch_flooding_f32(self, llr, syn=None, rule=RULE_SPA, n_ite=20, early_stop=True, norm=1.0, offset=0.0,
                           n_threads=0):
        llr = np.ascontiguousarray(llr, dtype=np.float32)
        F = llr.shape[0]
        syn = None if syn is None else np.ascontiguousarray(syn, dtype=np.uint8)
        post = np.zeros((F, self.N), dtype=np.float32)
        hard = np.zeros((F, self.N), dtype=np.uint8)
        iters = np.zeros(F, dtype=np.int32)
        ok = np.zeros(F, dtype=np.uint8)
        nt = lib().ora_batch_flooding_f32(self.h, _p(llr), _p(syn), F, rule, n_ite, int(early_stop), norm, offset,
                                          _p(post), _p(hard), _p(iters), _p(ok), n_threads)
        return hard, post, iters, ok.astype(bool), nt


def normalize_eighths(v, k):
    return lib().ora_normalize_eighths(int(v), int(k))


def pack_bits_msb(bits):
    """bits[..., n] (0/1) -> uint32[..., ceil(n/32)], MSB first inside every 32-bit word
    (errorcorrection/subcomponents/helpers.h:65-68 convention)."""
    bits = np.asarray(bits, dtype=np.uint8)
    n = bits.shape[-1]
    pad = (-n) % 32
    if pad:
        bits = np.concatenate([bits, np.zeros(bits.shape[:-1] + (pad,), np.uint8)], axis=-1)
    by = np.packbits(bits, axis=-1, bitorder="big")
    return by.reshape(bits.shape[:-1] + (-1, 4)).view(">u4").astype(np.uint32).reshape(bits.shape[:-1] + (-1,))


def unpack_bits_msb(words, n):
    words = np.asarray(words, dtype=np.uint32)
    by = words.astype(">u4").view(np.uint8).reshape(words.shape[:-1] + (-1,))
    return np.unpackbits(by, axis=-1, bitorder="big")[..., :n]
