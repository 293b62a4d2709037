"""Parity of the CUDA layered int8 path (through the C ABI) against the CPU oracle.
Bar: bit-exact decoded bits, syndrome-ok flag and iteration count."""
import numpy as np
import pytest

from conftest import make_frames

pytestmark = pytest.mark.gpu


def _run_case(q, O, data_dir, name, F, qber, mag, mode, rule, n_ite, early, offset=2, k8=6, out_all=True, seed=1,
              expect_kernel=None, random_llr=False, flags=0):
    path = "%s/%s" % (data_dir, name)
    oc = O.Code.from_qc(path)
    if random_llr:   # stress: arbitrary int8 LLRs incl. -128 / 127 saturation, random syndrome
        rng = np.random.default_rng(seed)
        llr = rng.integers(-128, 128, size=(F, oc.N)).astype(np.int32)
        syn = rng.integers(0, 2, size=(F, oc.M)).astype(np.uint8) if mode == "syndrome" else None
    else:
        llr, syn, _ = make_frames(O, oc, F, qber, mag, 31, mode, seed)
    code = q.Code.from_qc_file(path)
    dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=rule, dtype=q.DTYPE_I8, max_iter=n_ite, early_stop=early,
                    norm_factor=k8 / 8.0, offset=float(offset), out_mode=q.OUT_ALL if out_all else q.OUT_INFO, flags=flags)
    if expect_kernel is None:   # streamed kernel (layered_i8s.cu) for Z % 128 == 0, else the shared-memory-resident one
        streamed = oc.Z % 128 == 0 and not (flags & (q.FLAG_LI8_RESIDENT | q.FLAG_LI8_STREAM))
        expect_kernel = "layered_i8s_zpack4" if streamed else "layered_i8_zpack4"
    assert dec.kernel_name == expect_kernel, dec.kernel_name
    syn_packed = None if syn is None else q.pack_bits(syn)
    out, ok, iters, _ = dec.decode(llr.astype(np.int8), syn_packed)
    nbits = oc.N if out_all else oc.K
    got = q.unpack_bits(out, nbits)
    orule = O.RULE_NMS if rule == q.RULE_NMS else O.RULE_OMS
    hard, oit, ook, _ = oc.batch_layered_fixed_i8(llr.astype(np.int8), syn, rule=orule, n_ite=n_ite, early_stop=early,
                                                  offset=offset, norm_eighths=k8)
    bad_frames = np.nonzero((got != hard[:, :nbits]).any(axis=1))[0]
    assert bad_frames.size == 0, "bit mismatch in frames %s (first frame: %d differing bits, cols %s)" % (
        bad_frames[:8], (got[bad_frames[0]] != hard[bad_frames[0], :nbits]).sum(),
        np.unique(np.nonzero(got[bad_frames[0]] != hard[bad_frames[0], :nbits])[0] // oc.Z)[:16])
    assert (iters == oit).all(), "iteration mismatch: gpu %s oracle %s" % (iters[:16], oit[:16])
    assert (ok == ook).all()
    st = dec.stats()
    assert st["frames"] == F and st["iter_sum"] == int(oit.sum()) and st["failures"] == int((~ook).sum())
    assert st["kernel_launches"] >= 1
    dec.close()
    return iters, ok


@pytest.mark.parametrize("mode", ["parity", "syndrome"])
@pytest.mark.parametrize("rule,k8,offset", [("nms", 6, 0), ("oms", 8, 2), ("nms", 8, 0), ("nms", 7, 0)])
def test_bg1_z384_qber3(q, O, data_dir, mode, rule, k8, offset):
    r = q.RULE_NMS if rule == "nms" else q.RULE_OMS
    iters, ok = _run_case(q, O, data_dir, "NR_1_1_384.qc", 48, 0.03, 14, mode, r, 10, True, offset=offset, k8=k8)
    assert ok.all()


@pytest.mark.parametrize("mode", ["resident", "stream"])
def test_bg1_z384_both_message_placements(q, O, data_dir, mode):
    """the previous-generation kernel keeps check-to-variable messages either resident in shared memory (+ register rows,
    2 frames/SM) or streamed through an L2-resident scratch (4 frames/SM); both must be bit-exact"""
    fl = q.FLAG_LI8_RESIDENT if mode == "resident" else q.FLAG_LI8_STREAM
    _run_case(q, O, data_dir, "NR_1_1_384.qc", 40, 0.05, 12, "syndrome", q.RULE_NMS, 10, True, k8=6, seed=21, flags=fl)
    _run_case(q, O, data_dir, "NR_1_1_384.qc", 24, 0.03, 14, "parity", q.RULE_OMS, 3, False, offset=2, seed=22, flags=fl)
    _run_case(q, O, data_dir, "NR_2_3_112.qc", 24, 0.03, 14, "parity", q.RULE_NMS, 10, True, seed=23, flags=fl)


def test_bg1_z384_fixed_iterations_matlab_constants(q, O, data_dir):
    # ML/BPSK_nrldpc_sim_FP.m constants: offset 2, no early stop; 20 iterations cut to 6 for test time
    _run_case(q, O, data_dir, "NR_1_1_384.qc", 24, 0.06, 11, "syndrome", q.RULE_OMS, 6, False, offset=2)


def test_bg1_z384_info_only_output(q, O, data_dir):
    _run_case(q, O, data_dir, "NR_1_1_384.qc", 16, 0.03, 14, "parity", q.RULE_NMS, 10, True, out_all=False)


def test_bg1_z384_failing_frames_high_qber(q, O, data_dir):
    # far above threshold for pure syndrome mode: decoder must fail the same way the oracle does
    iters, ok = _run_case(q, O, data_dir, "NR_1_1_384.qc", 12, 0.20, 5, "syndrome", q.RULE_NMS, 5, True)
    assert not ok.all()


@pytest.mark.parametrize("mode", ["parity", "syndrome"])
def test_bg1_z384_random_saturating_llrs(q, O, data_dir, mode):
    # arbitrary int8 input incl. -128/127: exercises every clip in the datapath
    _run_case(q, O, data_dir, "NR_1_1_384.qc", 12, 0, 0, mode, q.RULE_OMS, 4, True, offset=1, random_llr=True)
    _run_case(q, O, data_dir, "NR_1_1_384.qc", 12, 0, 0, mode, q.RULE_NMS, 4, False, k8=5, random_llr=True, seed=7)


@pytest.mark.parametrize("name,mag", [("NR_1_1_192.qc", 14), ("NR_2_3_112.qc", 14), ("NR_1_1_24.qc", 14),
                                      ("NR_2_6_52.qc", 14), ("NR_1_7_240.qc", 14), ("NR_1_0_256.qc", 14)])
def test_other_lifting_sizes(q, O, data_dir, name, mag):
    # Z = 192, 112 (W not a multiple of 32), 24 and 52 (Z not a multiple of 32), 240, 256
    _run_case(q, O, data_dir, name, 20, 0.03, mag, "parity", q.RULE_NMS, 10, True)
    _run_case(q, O, data_dir, name, 20, 0.04, mag, "syndrome", q.RULE_OMS, 10, True, offset=1, seed=3)


@pytest.mark.parametrize("name", ["NR_2_1_384.qc", "NR_1_0_128.qc", "NR_2_0_256.qc"])
def test_streamed_kernel_other_base_graphs(q, O, data_dir, name):
    """the v5 streamed kernel (Z % 128 == 0) on BG2 and on the other lifting sizes: one, two and three warps per frame"""
    _run_case(q, O, data_dir, name, 40, 0.03, 14, "parity", q.RULE_NMS, 10, True, seed=31)
    _run_case(q, O, data_dir, name, 24, 0.05, 12, "syndrome", q.RULE_OMS, 8, True, offset=2, seed=32)
    _run_case(q, O, data_dir, name, 12, 0, 0, "syndrome", q.RULE_NMS, 3, False, k8=7, random_llr=True, seed=33)


def test_wifi_n1944(q, O, data_dir):
    # rate-1/2 N=1944 (Z=81 is not a multiple of 4 -> generic kernel)
    path = "%s/wifi_n1944_r12.qc" % data_dir
    oc = O.Code.from_qc(path)
    rng = np.random.default_rng(5)
    F = 16
    x = rng.integers(0, 2, (F, oc.N)).astype(np.uint8)
    e = (rng.random((F, oc.N)) < 0.04).astype(np.uint8)
    syn = np.stack([oc.syndrome(x[f]) for f in range(F)])
    llr = np.where(x ^ e, -12, 12).astype(np.int8)
    code = q.Code.from_qc_file(path)
    dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=20, early_stop=True,
                    norm_factor=0.75, out_mode=q.OUT_ALL)
    assert dec.kernel_name == "layered_generic"
    out, ok, iters, post = dec.decode(llr, q.pack_bits(syn), want_posterior=True)
    hard, oit, ook, _ = oc.batch_layered_fixed_i8(llr, syn, rule=O.RULE_NMS, n_ite=20, early_stop=True, norm_eighths=6)
    assert (q.unpack_bits(out, oc.N) == hard).all() and (iters == oit).all() and (ok == ook).all()
    for f in range(3):
        _, app, _, _ = oc.decode_layered_fixed(llr[f].astype(np.int32), syn[f], rule=O.RULE_NMS, n_ite=20, early_stop=True,
                                               norm_eighths=6)
        assert (post[f] == app).all()


def test_full_batch_roundtrip_property(q, O, data_dir):
    """BASELINE config-2 scale property test: encode -> BSC -> decode returns Alice's bits for every frame
    (size-independent; the oracle is only sampled)."""
    path = "%s/NR_1_1_384.qc" % data_dir
    oc = O.Code.from_qc(path)
    code = q.Code.from_qc_file(path)
    dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=10, early_stop=True,
                    norm_factor=0.75, out_mode=q.OUT_INFO)
    F = 4096
    rng = np.random.default_rng(11)
    msg = rng.integers(0, 2, (F, oc.K)).astype(np.uint8)
    cw = q.unpack_bits(dec.encode_nr(q.pack_bits(msg)), oc.N)
    assert (cw[:, :oc.K] == msg).all()
    for f in (0, 1, F - 1):   # GPU encoder against the oracle's restatement of nrldpc_encode.m
        assert (cw[f] == oc.nr_encode(msg[f])).all()
    syn = dec.syndrome(q.pack_bits(cw))
    assert not syn.any()
    noisy = cw.copy()
    noisy[:, :oc.K] ^= (rng.random((F, oc.K)) < 0.03).astype(np.uint8)
    known = np.zeros(oc.N, np.uint8)
    known[oc.K:] = 1
    llr = dec.make_llr(q.pack_bits(noisy), 14.0, 31.0, known_mask=q.pack_bits(known))
    assert llr.dtype == np.int8 and set(np.unique(llr)) <= {-31, -14, 14, 31}
    out, ok, iters, _ = dec.decode(llr)
    assert ok.all()
    assert (q.unpack_bits(out, oc.K) == msg).all()
    # the fused bits-in call (what the ecd2 handlers use) returns exactly the same
    out2, ok2, iters2 = dec.decode_bits(q.pack_bits(noisy), 14.0, 31.0, known_mask=q.pack_bits(known))
    assert (out2 == out).all() and (ok2 == ok).all() and (iters2 == iters).all()
    st = dec.stats()
    assert st["frames"] == 2 * F and st["failures"] == 0
    assert sum(st["iter_hist"]) == 2 * F and st["iter_sum"] == 2 * int(iters.sum())
    # sampled oracle agreement on iteration counts
    sel = rng.choice(F, 16, replace=False)
    hard, oit, ook, _ = oc.batch_layered_fixed_i8(llr[sel], None, rule=O.RULE_NMS, n_ite=10, early_stop=True, norm_eighths=6)
    assert (oit == iters[sel]).all() and ook.all()


def test_baseline_full_size_65536_frames_roundtrip(q, O, data_dir):
    """BASELINE config 2 at its full size (65 536 frames of BG1 Z=384 in ONE call through the host-pointer entry point):
    encode -> BSC(3 %) -> decode gives back every one of Alice's 553 648 128 bits; syndrome linearity over the whole batch."""
    path = "%s/NR_1_1_384.qc" % data_dir
    oc = O.Code.from_qc(path)
    code = q.Code.from_qc_file(path)
    dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=10, early_stop=True,
                    norm_factor=0.75, out_mode=q.OUT_INFO)
    F, C = 65536, 4096
    rng = np.random.default_rng(2024)
    kw = (oc.K + 31) // 32
    msg_p = np.empty((F, kw), np.uint32)
    noisy_p = np.empty((F, (oc.N + 31) // 32), np.uint32)
    for c in range(0, F, C):   # inputs are built chunk-wise on the host (numpy), the decode below is one call
        m = rng.integers(0, 2, (C, oc.K), dtype=np.uint8)
        msg_p[c:c + C] = q.pack_bits(m)
        cw = dec.encode_nr(msg_p[c:c + C])
        e = np.zeros((C, oc.N), np.uint8)
        e[:, :oc.K] = rng.random((C, oc.K), dtype=np.float32) < 0.03
        noisy_p[c:c + C] = cw ^ q.pack_bits(e)
    known = np.zeros(oc.N, np.uint8)
    known[oc.K:] = 1
    dec.reset_stats()
    out, ok, iters = dec.decode_bits(noisy_p, 14.0, 31.0, known_mask=q.pack_bits(known))
    assert ok.all()
    assert (out == msg_p).all()
    st = dec.stats()
    assert st["frames"] == F and st["failures"] == 0 and st["iter_sum"] == int(iters.sum())
    assert 1 <= iters.min() and iters.max() <= 4           # QBER 3 %: converges in 1-3 iterations
    # linearity of the syndrome over the full batch: H(a ^ b) = H(a) ^ H(b), with b = a cyclic shift of the batch
    sa = dec.syndrome(noisy_p)
    sb = np.roll(sa, 1, axis=0)
    assert (dec.syndrome(noisy_p ^ np.roll(noisy_p, 1, axis=0)) == (sa ^ sb)).all()
    # sampled oracle agreement on iteration counts at this size
    sel = rng.choice(F, 8, replace=False)
    llr = dec.make_llr(noisy_p[sel], 14.0, 31.0, known_mask=q.pack_bits(known))
    _, oit, ook, _ = oc.batch_layered_fixed_i8(llr, None, rule=O.RULE_NMS, n_ite=10, early_stop=True, norm_eighths=6)
    assert (oit == iters[sel]).all() and ook.all()


def test_host_pipeline_chunks_on_two_streams(q, O, data_dir):
    """the host-pointer entry points cut the batch into chunks (2, 4, 8 ... waves of the persistent grid) that ping-pong over
    two streams; kernels of the two streams overlap in time, so nothing they write may be shared (regression:
    streamed-message scratch, and with bit input the slots' extension-column scratch)"""
    path = "%s/NR_1_1_384.qc" % data_dir
    oc = O.Code.from_qc(path)
    code = q.Code.from_qc_file(path)
    dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=10, early_stop=True,
                    norm_factor=0.75, out_mode=q.OUT_INFO)
    F = 9000
    rng = np.random.default_rng(31)
    msg = rng.integers(0, 2, (F, oc.K)).astype(np.uint8)
    cw = q.unpack_bits(dec.encode_nr(q.pack_bits(msg)), oc.N)
    noisy = cw.copy()
    noisy[:, :oc.K] ^= (rng.random((F, oc.K)) < 0.045).astype(np.uint8)     # ~2-3 iterations: every row's messages are re-read
    known = np.zeros(oc.N, np.uint8)
    known[oc.K:] = 1
    for rep in range(3):
        out, ok, iters = dec.decode_bits(q.pack_bits(noisy), 12.0, 31.0, known_mask=q.pack_bits(known))
        assert ok.all() and (q.unpack_bits(out, oc.K) == msg).all(), "rep %d: %d frames wrong" % (
            rep, int((q.unpack_bits(out, oc.K) != msg).any(axis=1).sum()))
    llr = dec.make_llr(q.pack_bits(noisy), 12.0, 31.0, known_mask=q.pack_bits(known))
    out2, ok2, iters2, _ = dec.decode(llr)
    assert (out2 == out).all() and (iters2 == iters).all()
    sel = rng.choice(F, 24, replace=False)
    hard, oit, ook, _ = oc.batch_layered_fixed_i8(llr[sel], None, rule=O.RULE_NMS, n_ite=10, early_stop=True, norm_eighths=6)
    assert (oit == iters[sel]).all() and (hard[:, :oc.K] == msg[sel]).all()


@pytest.mark.parametrize("mode,qber,mag", [("parity", 0.03, 14), ("syndrome", 0.06, 11)])
def test_bg1_z384_many_frames_per_slot_vs_oracle(q, O, data_dir, mode, qber, mag):
    """more frames than the persistent grid has slots (148 SMs x 5): every slot switches frames several times (frame prefetch,
    barrier parities and message scratch carried across frames) and groups of one CTA finish at different iterations;
    bits, iteration counts and ok flags still equal the oracle's, frame by frame"""
    iters, ok = _run_case(q, O, data_dir, "NR_1_1_384.qc", 2600, qber, mag, mode, q.RULE_NMS, 10, True, out_all=False, seed=91)
    assert len(set(iters.tolist())) >= 2          # a mix of iteration counts inside the batch


@pytest.mark.parametrize("name,mode,qber,mag", [("NR_1_1_384.qc", "parity", 0.03, 14), ("NR_1_1_384.qc", "syndrome", 0.07, 11),
                                                ("NR_2_0_256.qc", "parity", 0.04, 12)])
def test_discard_scratch_flag_keeps_results(q, O, data_dir, name, mode, qber, mag):
    """QLDPC_FLAG_DISCARD_SCRATCH drops the L2 lines of a slot's message scratch when its frame ends (discard.global.L2); the
    slot's next frame rewrites every line before it reads one, so bits / ok / iteration counts still equal the oracle's --
    checked with several frames per slot and frames that need many iterations (messages re-read after every iteration)"""
    iters, ok = _run_case(q, O, data_dir, name, 2400, qber, mag, mode, q.RULE_NMS, 10, True, out_all=False, seed=17,
                          flags=q.FLAG_DISCARD_SCRATCH)
    if name == "NR_1_1_384.qc":
        assert iters.max() >= 2                    # messages written, re-read, discarded and rewritten by the slot's next frame


@pytest.mark.parametrize("mode", ["parity", "syndrome"])
@pytest.mark.parametrize("name", ["NR_1_1_384.qc", "NR_2_0_256.qc", "NR_1_0_128.qc"])
def test_bit_input_fused_synthesis_equals_two_kernel_path(q, O, data_dir, name, mode):
    """qldpc_decode_bits on the streamed kernel synthesises the LLRs INSIDE the decoder (bit input); bits, flags and
    iteration counts must equal the two-kernel path (QLDPC_FLAG_NO_FUSED_BITS: make_llr + decode) and the oracle on the
    same LLRs -- with known, punctured and noisy positions mixed inside core AND extension columns, more frames than slots"""
    path = "%s/%s" % (data_dir, name)
    oc = O.Code.from_qc(path)
    code = q.Code.from_qc_file(path)
    F = 2100
    rng = np.random.default_rng(77)
    x = rng.integers(0, 2, (F, oc.N)).astype(np.uint8)
    kw = dict(schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=10, early_stop=True, norm_factor=0.75,
              out_mode=q.OUT_ALL)
    fused, plain = q.Decoder(code, **kw), q.Decoder(code, flags=q.FLAG_NO_FUSED_BITS, **kw)
    assert fused.kernel_name == plain.kernel_name == "layered_i8s_zpack4"
    known = np.zeros(oc.N, np.uint8)
    punct = np.zeros(oc.N, np.uint8)
    if mode == "parity":
        cw = q.unpack_bits(fused.encode_nr(q.pack_bits(x[:, :oc.K])), oc.N)
        known[oc.K:] = 1
        known[rng.choice(oc.K, oc.K // 50, replace=False)] = 1        # some revealed information bits (blind reconciliation)
        punct[oc.N - 3 * oc.Z // 2:] = 1                               # punctured tail, not a whole number of block columns
        known[punct == 1] = 0
        syn = None
    else:
        cw = x
        punct[rng.choice(oc.N, oc.N // 40, replace=False)] = 1
        syn = fused.syndrome(q.pack_bits(cw))
    noise = (rng.random((F, oc.N)) < 0.03).astype(np.uint8) & (1 - known)
    noisy_p = q.pack_bits(cw ^ noise)
    km, pm = q.pack_bits(known), q.pack_bits(punct)
    a = fused.decode_bits(noisy_p, 13.0, 31.0, known_mask=km, punct_mask=pm, syndrome=syn)
    b = plain.decode_bits(noisy_p, 13.0, 31.0, known_mask=km, punct_mask=pm, syndrome=syn)
    for u, v in zip(a, b):
        assert (u == v).all()
    sel = rng.choice(F, 40, replace=False)
    llr = plain.make_llr(noisy_p[sel], 13.0, 31.0, known_mask=km, punct_mask=pm)
    hard, oit, ook, _ = oc.batch_layered_fixed_i8(llr, None if syn is None else q.unpack_bits(syn[sel], oc.M), rule=O.RULE_NMS,
                                                  n_ite=10, early_stop=True, norm_eighths=6)
    assert (q.unpack_bits(a[0][sel], oc.N) == hard).all() and (a[2][sel] == oit).all() and (a[1][sel] == ook).all()
    fused.close()
    plain.close()


def test_shortened_prefix_info_bits_on_the_int8_kernels(q, O, data_dir):
    """information positions [0, k') with k' smaller than the systematic part (a shortened code) -- k' a whole
    number of block columns is written directly, any other k' goes through the gather; frames must not be mis-strided"""
    path = "%s/NR_1_1_384.qc" % data_dir
    oc = O.Code.from_qc(path)
    llr, syn, truth = make_frames(O, oc, 900, 0.03, 14, 31, "parity", 5)
    for kk in (20 * oc.Z, 20 * oc.Z + 100, 37):
        code = q.Code.from_qc_file(path)
        code.set_info_bits_pos(np.arange(kk))
        dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=10, early_stop=True,
                        norm_factor=0.75, out_mode=q.OUT_INFO)
        assert dec.out_words == (kk + 31) // 32
        out, ok, iters, _ = dec.decode(llr.astype(np.int8))
        assert ok.all() and (q.unpack_bits(out, kk) == truth[:, :kk]).all()
        dec.close()


def test_llr_magnitudes_outside_the_dtype_range_are_rejected(q, data_dir):
    """llr_known = 200 on the int8 tier used to wrap to -56 and flip every confirmed bit"""
    code = q.Code.from_qc_file("%s/NR_1_1_384.qc" % data_dir)
    dec = q.Decoder(code, schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=2, norm_factor=0.75)
    bits = np.zeros((4, dec.cw_words), np.uint32)
    for noisy, known in ((14.0, 200.0), (-1.0, 31.0), (128.0, 31.0)):
        with pytest.raises(q.QldpcError) as e:
            dec.make_llr(bits, noisy, known)
        assert e.value.code == 1
        with pytest.raises(q.QldpcError):
            dec.decode_bits(bits, noisy, known)
    assert (dec.make_llr(bits, 127.0, 127.0) == 127).all()
    dec.close()


@pytest.mark.parametrize("with_syndrome", [False, True])
def test_zero_copy_decode_bits_on_pinned_host_buffers(q, O, data_dir, with_syndrome):
    """qldpc_decode_bits with PINNED host buffers: the decoder kernel reads the packed bits and writes bits / ok / iteration
    counts in place over PCIe (one launch, no staging copies); results equal the chunked copy pipeline (pageable numpy
    buffers, and pinned buffers with QLDPC_FLAG_NO_ZERO_COPY) and the oracle"""
    import torch
    path = "%s/NR_1_1_384.qc" % data_dir
    oc = O.Code.from_qc(path)
    code = q.Code.from_qc_file(path)
    kw = dict(schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=10, early_stop=True, norm_factor=0.75,
              out_mode=q.OUT_INFO)
    dec, staged = q.Decoder(code, **kw), q.Decoder(code, flags=q.FLAG_NO_ZERO_COPY, **kw)
    F = 4000
    rng = np.random.default_rng(11)
    known = np.zeros(oc.N, np.uint8)
    if with_syndrome:
        x = rng.integers(0, 2, (F, oc.N)).astype(np.uint8)
        syn = dec.syndrome(q.pack_bits(x))
        noisy = x ^ (rng.random((F, oc.N)) < 0.04).astype(np.uint8)
        want = x[:, :oc.K]
    else:
        msg = rng.integers(0, 2, (F, oc.K)).astype(np.uint8)
        noisy = q.unpack_bits(dec.encode_nr(q.pack_bits(msg)), oc.N)
        noisy[:, :oc.K] ^= (rng.random((F, oc.K)) < 0.04).astype(np.uint8)
        known[oc.K:] = 1
        syn, want = None, msg
    bits_np, km = q.pack_bits(noisy), q.pack_bits(known)
    ref = dec.decode_bits(bits_np, 12.0, 31.0, known_mask=km, syndrome=syn)          # pageable: chunked pipeline
    launches0 = dec.stats()["kernel_launches"]
    h_bits = torch.from_numpy(bits_np.view(np.int32)).pin_memory()
    h_syn = None if syn is None else torch.from_numpy(syn.view(np.int32)).pin_memory()
    h_out = torch.zeros((F, dec.out_words), dtype=torch.int32).pin_memory()
    h_ok = torch.zeros(F, dtype=torch.uint8).pin_memory()
    h_it = torch.zeros(F, dtype=torch.int16).pin_memory()
    for d in (dec, staged):
        h_out.zero_(); h_ok.zero_(); h_it.zero_()
        rc = q.lib().qldpc_decode_bits(d.h, h_bits.data_ptr(), km.ctypes.data, None, 12.0, 31.0,
                                       None if h_syn is None else h_syn.data_ptr(), F, h_out.data_ptr(), h_ok.data_ptr(),
                                       h_it.data_ptr())
        assert rc == 0
        assert (h_out.numpy().view(np.uint32) == ref[0]).all() and (h_ok.numpy().astype(bool) == ref[1]).all()
        assert (h_it.numpy().view(np.uint16) == ref[2]).all()
    assert dec.stats()["kernel_launches"] - launches0 == 2          # magnitude table + ONE decode launch
    assert ref[1].mean() > 0.95 and (q.unpack_bits(ref[0], oc.K)[ref[1]] == want[ref[1]]).all()
    sel = rng.choice(F, 32, replace=False)
    llr = dec.make_llr(bits_np[sel], 12.0, 31.0, known_mask=km)
    hard, oit, ook, _ = oc.batch_layered_fixed_i8(llr, None if syn is None else q.unpack_bits(syn[sel], oc.M), rule=O.RULE_NMS,
                                                  n_ite=10, early_stop=True, norm_eighths=6)
    assert (oit == ref[2][sel]).all() and (ook == ref[1][sel]).all() and (hard[:, :oc.K] == q.unpack_bits(ref[0][sel], oc.K)).all()
    dec.close()
    staged.close()
