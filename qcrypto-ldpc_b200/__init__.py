"""Python binding of libqldpc_b200.so (C ABI: include/qldpc.h).

Thin ctypes layer used by the tests and bench.py; the product is the shared library.  Import this
package with importlib (the directory name carries a dash):

    import importlib; q = importlib.import_module("qcrypto-ldpc_b200")

There is NO CPU fallback here: if the CUDA library is missing or no sm_100 device is visible every
compute call raises QldpcError.  The CPU oracle under oracle/ is test infrastructure and is never
imported from this package.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# QLDPC_LIB overrides the library path (kernel experiments build variant libraries side by side)
LIB_PATH = os.environ.get("QLDPC_LIB") or os.path.join(_HERE, "libqldpc_b200.so")
DATA_DIR = os.path.join(_HERE, "data")

SCHED_FLOODING, SCHED_LAYERED = 0, 1
RULE_SPA, RULE_NMS, RULE_OMS = 0, 1, 2
DTYPE_F32, DTYPE_I16, DTYPE_I8 = 0, 1, 2
OUT_INFO, OUT_ALL = 0, 1
ITER_HIST_BINS = 64
MAX_DEVICES = 8
FLAG_L2_PERSIST, FLAG_LI8_RESIDENT, FLAG_LI8_STREAM, FLAG_NO_FUSED_BITS, FLAG_NO_ZERO_COPY, FLAG_FAST_SPA = 1, 2, 4, 8, 16, 32
FLAG_DISCARD_SCRATCH = 64

_NP_DTYPE = {DTYPE_F32: np.float32, DTYPE_I16: np.int16, DTYPE_I8: np.int8}

# every symbol include/qldpc.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "qldpc_code_from_alist_file", "qldpc_code_from_qc_file", "qldpc_code_from_qc", "qldpc_code_from_csr",
    "qldpc_code_set_info_bits_pos", "qldpc_code_get_info", "qldpc_code_free",
    "qldpc_decoder_config_default", "qldpc_decoder_create", "qldpc_decoder_free",
    "qldpc_out_words", "qldpc_syndrome_words", "qldpc_codeword_words",
    "qldpc_decode", "qldpc_decode_device", "qldpc_syndrome", "qldpc_syndrome_device",
    "qldpc_make_llr", "qldpc_make_llr_device", "qldpc_decode_bits", "qldpc_decode_bits_device", "qldpc_encode_nr", "qldpc_encode_nr_device",
    "qldpc_get_stats", "qldpc_reset_stats", "qldpc_decoder_kernel_name", "qldpc_strerror",
    "qldpc_last_cuda_error", "qldpc_version", "qldpc_privacy_amplify", "qldpc_crc32_frames",
    "qldpc_encoder_from_h", "qldpc_encoder_from_g_alist_file", "qldpc_encoder_get_info", "qldpc_encoder_info_bits_pos",
    "qldpc_encode", "qldpc_encode_device", "qldpc_encoder_free",
]


class QldpcError(RuntimeError):
    def __init__(self, code, where=""):
        self.code = code
        msg = lib().qldpc_strerror(code).decode()
        cuda = lib().qldpc_last_cuda_error().decode()
        super().__init__("%s: %s (code %d)%s" % (where, msg, code, (" [" + cuda + "]") if cuda and code == 5 else ""))


class CodeInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n", "m", "k", "edges", "z", "base_rows", "base_cols", "max_chk_degree",
                                         "max_var_degree")]


class DecoderConfig(C.Structure):
    _fields_ = [("schedule", C.c_int32), ("rule", C.c_int32), ("dtype", C.c_int32), ("max_iter", C.c_int32),
                ("early_stop", C.c_int32), ("syndrome_depth", C.c_int32), ("norm_factor", C.c_float),
                ("offset", C.c_float), ("msg_max", C.c_int32), ("app_max", C.c_int32), ("out_mode", C.c_int32),
                ("device", C.c_int32), ("n_devices", C.c_int32), ("devices", C.c_int32 * MAX_DEVICES), ("flags", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("failures", C.c_uint64), ("iter_sum", C.c_uint64),
                ("iter_hist", C.c_uint64 * ITER_HIST_BINS), ("kernel_launches", C.c_uint64)]


_lib = None


def lib():
    """Loads libqldpc_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libqldpc_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "or `make -C qcrypto-ldpc_b200/csrc`")
        L = C.CDLL(LIB_PATH)
        vp, i32, pp = C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)
        L.qldpc_code_from_alist_file.argtypes = [C.c_char_p, pp]
        L.qldpc_code_from_qc_file.argtypes = [C.c_char_p, pp]
        L.qldpc_code_from_qc.argtypes = [vp, i32, i32, i32, pp]
        L.qldpc_code_from_csr.argtypes = [i32, i32, vp, vp, pp]
        L.qldpc_code_set_info_bits_pos.argtypes = [vp, vp, i32]
        L.qldpc_code_get_info.argtypes = [vp, C.POINTER(CodeInfo)]
        L.qldpc_code_free.argtypes = [vp]
        L.qldpc_code_free.restype = None
        L.qldpc_decoder_config_default.argtypes = [C.POINTER(DecoderConfig)]
        L.qldpc_decoder_config_default.restype = None
        L.qldpc_decoder_create.argtypes = [vp, C.POINTER(DecoderConfig), pp]
        L.qldpc_decoder_free.argtypes = [vp]
        L.qldpc_decoder_free.restype = None
        for n in ("qldpc_out_words", "qldpc_syndrome_words", "qldpc_codeword_words"):
            getattr(L, n).argtypes = [vp]
            getattr(L, n).restype = i32
        L.qldpc_decode.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp]
        L.qldpc_decode_device.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp]
        L.qldpc_syndrome.argtypes = [vp, vp, i32, vp]
        L.qldpc_syndrome_device.argtypes = [vp, vp, i32, vp, vp]
        L.qldpc_make_llr.argtypes = [vp, vp, vp, vp, C.c_float, C.c_float, i32, vp]
        L.qldpc_make_llr_device.argtypes = [vp, vp, vp, vp, C.c_float, C.c_float, i32, vp, vp]
        L.qldpc_decode_bits.argtypes = [vp, vp, vp, vp, C.c_float, C.c_float, vp, i32, vp, vp, vp]
        L.qldpc_decode_bits_device.argtypes = [vp, vp, vp, vp, C.c_float, C.c_float, vp, i32, vp, vp, vp, vp]
        L.qldpc_encode_nr.argtypes = [vp, vp, i32, vp]
        L.qldpc_encode_nr_device.argtypes = [vp, vp, i32, vp, vp]
        L.qldpc_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.qldpc_reset_stats.argtypes = [vp]
        L.qldpc_decoder_kernel_name.argtypes = [vp]
        L.qldpc_decoder_kernel_name.restype = C.c_char_p
        L.qldpc_strerror.argtypes = [C.c_int]
        L.qldpc_strerror.restype = C.c_char_p
        L.qldpc_last_cuda_error.restype = C.c_char_p
        L.qldpc_version.restype = C.c_int
        L.qldpc_encoder_from_h.argtypes = [vp, i32, pp]
        L.qldpc_encoder_from_g_alist_file.argtypes = [C.c_char_p, i32, pp]
        L.qldpc_encoder_get_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
        L.qldpc_encoder_info_bits_pos.argtypes = [vp, vp]
        L.qldpc_encode.argtypes = [vp, vp, i32, vp]
        L.qldpc_encode_device.argtypes = [vp, vp, i32, vp, vp]
        L.qldpc_encoder_free.argtypes = [vp]
        L.qldpc_encoder_free.restype = None
        L.qldpc_privacy_amplify.argtypes = [i32, vp, i32, vp, vp, vp, i32, vp, i32]
        L.qldpc_crc32_frames.argtypes = [i32, vp, i32, i32, i32, vp]
        _lib = L
    return _lib


def _chk(rc, where):
    if rc != 0:
        raise QldpcError(rc, where)


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def data_path(name):
    return os.path.join(DATA_DIR, name)


class Code:
    """A parity-check matrix (host side).  Mirrors what the reference drivers hand to the AFF3CT
    decoder constructor: H plus info_bits_pos (BOOT/src/main.cpp:175-193)."""

    def __init__(self, handle):
        self.h = handle
        info = CodeInfo()
        _chk(lib().qldpc_code_get_info(self.h, C.byref(info)), "qldpc_code_get_info")
        self.info = info
        for f, _ in CodeInfo._fields_:
            setattr(self, f, getattr(info, f))

    @classmethod
    def from_alist(cls, path):
        h = C.c_void_p()
        _chk(lib().qldpc_code_from_alist_file(str(path).encode(), C.byref(h)), "qldpc_code_from_alist_file")
        return cls(h)

    @classmethod
    def from_qc_file(cls, path):
        h = C.c_void_p()
        _chk(lib().qldpc_code_from_qc_file(str(path).encode(), C.byref(h)), "qldpc_code_from_qc_file")
        return cls(h)

    @classmethod
    def from_qc(cls, base, z):
        base = np.ascontiguousarray(base, dtype=np.int32)
        h = C.c_void_p()
        _chk(lib().qldpc_code_from_qc(_np_ptr(base), base.shape[0], base.shape[1], int(z), C.byref(h)), "qldpc_code_from_qc")
        return cls(h)

    @classmethod
    def from_csr(cls, n, m, row_ptr, col_idx):
        rp = np.ascontiguousarray(row_ptr, dtype=np.int32)
        ci = np.ascontiguousarray(col_idx, dtype=np.int32)
        h = C.c_void_p()
        _chk(lib().qldpc_code_from_csr(n, m, _np_ptr(rp), _np_ptr(ci), C.byref(h)), "qldpc_code_from_csr")
        return cls(h)

    def set_info_bits_pos(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.int32)
        _chk(lib().qldpc_code_set_info_bits_pos(self.h, _np_ptr(pos), len(pos)), "qldpc_code_set_info_bits_pos")
        self.__init__(self.h)

    def __del__(self):
        try:
            if self.h:
                lib().qldpc_code_free(self.h)
                self.h = None
        except Exception:
            pass


class Decoder:
    """Batched decoder bound to one CUDA device (qldpc_decoder_create)."""

    def __init__(self, code, schedule=SCHED_FLOODING, rule=RULE_SPA, dtype=DTYPE_F32, max_iter=100, early_stop=True,
                 syndrome_depth=1, norm_factor=1.0, offset=0.0, msg_max=0, app_max=0, out_mode=OUT_INFO, device=0,
                 devices=None, flags=0):
        """devices: list of CUDA ordinals -> one multi-device decoder whose host-buffer calls shard the frames over them
        (qldpc_decoder_config.devices); flags: FLAG_* bits"""
        cfg = DecoderConfig()
        lib().qldpc_decoder_config_default(C.byref(cfg))
        cfg.schedule, cfg.rule, cfg.dtype, cfg.max_iter = schedule, rule, dtype, max_iter
        cfg.early_stop, cfg.syndrome_depth = int(early_stop), syndrome_depth
        cfg.norm_factor, cfg.offset, cfg.msg_max, cfg.app_max = norm_factor, offset, msg_max, app_max
        cfg.out_mode, cfg.device = out_mode, device
        cfg.flags = flags
        if devices is not None:
            cfg.n_devices = len(devices)
            for k, dv in enumerate(devices):
                cfg.devices[k] = dv
        self.cfg = cfg
        self.code = code
        self.h = C.c_void_p()
        _chk(lib().qldpc_decoder_create(code.h, C.byref(cfg), C.byref(self.h)), "qldpc_decoder_create")
        self.out_words = lib().qldpc_out_words(self.h)
        self.syn_words = lib().qldpc_syndrome_words(self.h)
        self.cw_words = lib().qldpc_codeword_words(self.h)
        self.kernel_name = lib().qldpc_decoder_kernel_name(self.h).decode()
        self.np_dtype = _NP_DTYPE[dtype]

    def close(self):
        if self.h:
            lib().qldpc_decoder_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host-buffer entry points (numpy arrays; copies happen inside the library)
    def decode(self, llr, syndrome=None, want_posterior=False):
        llr = np.ascontiguousarray(llr, dtype=self.np_dtype)
        F = llr.shape[0]
        assert llr.shape[1] == self.code.n
        syn = None if syndrome is None else np.ascontiguousarray(syndrome, dtype=np.uint32)
        out = np.zeros((F, self.out_words), dtype=np.uint32)
        ok = np.zeros(F, dtype=np.uint8)
        iters = np.zeros(F, dtype=np.uint16)
        post = None
        if want_posterior:
            post = np.zeros((F, self.code.n), dtype=np.float32 if self.cfg.dtype == DTYPE_F32 else np.int32)
        _chk(lib().qldpc_decode(self.h, _np_ptr(llr), _np_ptr(syn), F, _np_ptr(out), _np_ptr(ok), _np_ptr(iters),
                                _np_ptr(post)), "qldpc_decode")
        return out, ok.astype(bool), iters, post

    def syndrome(self, bits_packed):
        bits = np.ascontiguousarray(bits_packed, dtype=np.uint32)
        F = bits.shape[0]
        syn = np.zeros((F, self.syn_words), dtype=np.uint32)
        _chk(lib().qldpc_syndrome(self.h, _np_ptr(bits), F, _np_ptr(syn)), "qldpc_syndrome")
        return syn

    def make_llr(self, bits_packed, llr_noisy, llr_known=0.0, known_mask=None, punct_mask=None):
        bits = np.ascontiguousarray(bits_packed, dtype=np.uint32)
        F = bits.shape[0]
        km = None if known_mask is None else np.ascontiguousarray(known_mask, dtype=np.uint32)
        pm = None if punct_mask is None else np.ascontiguousarray(punct_mask, dtype=np.uint32)
        out = np.zeros((F, self.code.n), dtype=self.np_dtype)
        _chk(lib().qldpc_make_llr(self.h, _np_ptr(bits), _np_ptr(km), _np_ptr(pm), llr_noisy, llr_known, F, _np_ptr(out)),
             "qldpc_make_llr")
        return out

    def decode_bits(self, bits_packed, llr_noisy, llr_known=0.0, known_mask=None, punct_mask=None, syndrome=None):
        """fused LLR synthesis + decode from packed sifted-key bits (the ecd2-facing call)"""
        bits = np.ascontiguousarray(bits_packed, dtype=np.uint32)
        F = bits.shape[0]
        km = None if known_mask is None else np.ascontiguousarray(known_mask, dtype=np.uint32)
        pm = None if punct_mask is None else np.ascontiguousarray(punct_mask, dtype=np.uint32)
        syn = None if syndrome is None else np.ascontiguousarray(syndrome, dtype=np.uint32)
        out = np.zeros((F, self.out_words), dtype=np.uint32)
        ok = np.zeros(F, dtype=np.uint8)
        iters = np.zeros(F, dtype=np.uint16)
        _chk(lib().qldpc_decode_bits(self.h, _np_ptr(bits), _np_ptr(km), _np_ptr(pm), llr_noisy, llr_known, _np_ptr(syn), F,
                                     _np_ptr(out), _np_ptr(ok), _np_ptr(iters)), "qldpc_decode_bits")
        return out, ok.astype(bool), iters

    def encode_nr(self, msg_packed):
        msg = np.ascontiguousarray(msg_packed, dtype=np.uint32)
        F = msg.shape[0]
        cw = np.zeros((F, self.cw_words), dtype=np.uint32)
        _chk(lib().qldpc_encode_nr(self.h, _np_ptr(msg), F, _np_ptr(cw)), "qldpc_encode_nr")
        return cw

    # ---- device-pointer entry points (raw addresses, e.g. torch.Tensor.data_ptr())
    def decode_device(self, d_llr, d_syndrome, n_frames, d_out, d_ok=0, d_iters=0, d_posterior=0, stream=0):
        _chk(lib().qldpc_decode_device(self.h, d_llr, d_syndrome or None, n_frames, d_out, d_ok or None, d_iters or None,
                                       d_posterior or None, stream or None), "qldpc_decode_device")

    def syndrome_device(self, d_bits, n_frames, d_syn, stream=0):
        _chk(lib().qldpc_syndrome_device(self.h, d_bits, n_frames, d_syn, stream or None), "qldpc_syndrome_device")

    def make_llr_device(self, d_bits, d_known, d_punct, llr_noisy, llr_known, n_frames, d_out, stream=0):
        _chk(lib().qldpc_make_llr_device(self.h, d_bits, d_known or None, d_punct or None, llr_noisy, llr_known, n_frames,
                                         d_out, stream or None), "qldpc_make_llr_device")

    def encode_nr_device(self, d_msg, n_frames, d_cw, stream=0):
        _chk(lib().qldpc_encode_nr_device(self.h, d_msg, n_frames, d_cw, stream or None), "qldpc_encode_nr_device")

    def stats(self):
        s = Stats()
        _chk(lib().qldpc_get_stats(self.h, C.byref(s)), "qldpc_get_stats")
        return {"frames": s.frames, "failures": s.failures, "iter_sum": s.iter_sum,
                "iter_hist": list(s.iter_hist), "kernel_launches": s.kernel_launches}

    def reset_stats(self):
        _chk(lib().qldpc_reset_stats(self.h), "qldpc_reset_stats")


class Encoder:
    """Systematic encoder of a code without the NR structure: Encoder.from_H(code) (Encoder_LDPC_from_H) or
    Encoder.from_G_alist(path) (Encoder_LDPC with a generator-matrix file)."""

    def __init__(self, handle):
        self.h = handle
        k, n = C.c_int32(), C.c_int32()
        _chk(lib().qldpc_encoder_get_info(self.h, C.byref(k), C.byref(n)), "qldpc_encoder_get_info")
        self.k, self.n = k.value, n.value
        pos = np.zeros(self.k, dtype=np.int32)
        self.info_bits_pos = pos if lib().qldpc_encoder_info_bits_pos(self.h, _np_ptr(pos)) == 0 else None

    @classmethod
    def from_H(cls, code, device=0):
        h = C.c_void_p()
        _chk(lib().qldpc_encoder_from_h(code.h, device, C.byref(h)), "qldpc_encoder_from_h")
        return cls(h)

    @classmethod
    def from_G_alist(cls, path, device=0):
        h = C.c_void_p()
        _chk(lib().qldpc_encoder_from_g_alist_file(str(path).encode(), device, C.byref(h)), "qldpc_encoder_from_g_alist_file")
        return cls(h)

    def encode(self, msg_packed):
        msg = np.ascontiguousarray(msg_packed, dtype=np.uint32)
        F = msg.shape[0]
        cw = np.zeros((F, (self.n + 31) // 32), dtype=np.uint32)
        _chk(lib().qldpc_encode(self.h, _np_ptr(msg), F, _np_ptr(cw)), "qldpc_encode")
        return cw

    def close(self):
        if self.h:
            lib().qldpc_encoder_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_bits(bits):
    """bits[..., n] of 0/1 -> uint32[..., ceil(n/32)], MSB-first in every word
    (errorcorrection/subcomponents/helpers.h:65-68)."""
    bits = np.asarray(bits, dtype=np.uint8)
    n = bits.shape[-1]
    pad = (-n) % 32
    if pad:
        bits = np.concatenate([bits, np.zeros(bits.shape[:-1] + (pad,), np.uint8)], axis=-1)
    by = np.packbits(bits, axis=-1, bitorder="big")
    return np.ascontiguousarray(by).view(">u4").astype(np.uint32)


def unpack_bits(words, n):
    words = np.ascontiguousarray(words, dtype=np.uint32)
    by = words.astype(">u4").view(np.uint8)
    return np.unpackbits(by, axis=-1, bitorder="big")[..., :n]


def privacy_amplify(keys, workbits, final_bits, seeds, device=0):
    """qldpc_privacy_amplify on a batch of blocks: keys [B, words] uint32 (MSB-first), per-block workbits / final_bits / seeds.
    Returns [B, ceil(max(final_bits)/32)] uint32 (words beyond a block's own length are zero)."""
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    B = keys.shape[0]
    wb = np.ascontiguousarray(workbits, dtype=np.int32)
    fb = np.ascontiguousarray(final_bits, dtype=np.int32)
    sd = np.ascontiguousarray(seeds, dtype=np.uint32)
    ow = max(1, (int(fb.max()) + 31) // 32) if B else 1
    out = np.zeros((B, ow), dtype=np.uint32)
    _chk(lib().qldpc_privacy_amplify(device, _np_ptr(keys), keys.shape[1], _np_ptr(wb), _np_ptr(fb), _np_ptr(sd), B, _np_ptr(out), ow),
         "qldpc_privacy_amplify")
    return out


def crc32_frames(bits_packed, words_per_frame=None, device=0):
    bits = np.ascontiguousarray(bits_packed, dtype=np.uint32)
    F, stride = bits.shape
    wpf = stride if words_per_frame is None else words_per_frame
    out = np.zeros(F, dtype=np.uint32)
    _chk(lib().qldpc_crc32_frames(device, _np_ptr(bits), F, wpf, stride, _np_ptr(out)), "qldpc_crc32_frames")
    return out
