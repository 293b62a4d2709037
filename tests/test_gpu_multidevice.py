"""Frame sharding INSIDE the library (qldpc_decoder_config.devices): one decoder handle, host-pointer calls cut
the batch into contiguous ranges, one host thread + stream pair per device, statistics summed on the host."""
import numpy as np
import pytest
import torch

from conftest import make_frames

pytestmark = pytest.mark.gpu


def _case(q, O, data_dir, devices):
    path = "%s/NR_1_1_384.qc" % data_dir
    oc = O.Code.from_qc(path)
    code = q.Code.from_qc_file(path)
    kw = dict(schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=10, early_stop=True, norm_factor=0.75,
              out_mode=q.OUT_INFO)
    multi = q.Decoder(code, devices=devices, **kw)
    single = q.Decoder(code, device=devices[0], **kw)
    F = 3001                                                   # ragged split
    rng = np.random.default_rng(8)
    msg = rng.integers(0, 2, (F, oc.K)).astype(np.uint8)
    cw_p = multi.encode_nr(q.pack_bits(msg))
    assert (cw_p == single.encode_nr(q.pack_bits(msg))).all()
    noisy = q.unpack_bits(cw_p, oc.N)
    noisy[:, :oc.K] ^= (rng.random((F, oc.K)) < 0.04).astype(np.uint8)
    known = np.zeros(oc.N, np.uint8)
    known[oc.K:] = 1
    a = multi.decode_bits(q.pack_bits(noisy), 13.0, 31.0, known_mask=q.pack_bits(known))
    b = single.decode_bits(q.pack_bits(noisy), 13.0, 31.0, known_mask=q.pack_bits(known))
    for u, v in zip(a, b):
        assert (u == v).all()
    assert a[1].all() and (q.unpack_bits(a[0], oc.K) == msg).all()
    llr = multi.make_llr(q.pack_bits(noisy), 13.0, 31.0, known_mask=q.pack_bits(known))
    c = multi.decode(llr)
    assert (c[0] == a[0]).all() and (c[2] == a[2]).all()
    assert (multi.syndrome(cw_p) == 0).all()
    sm, ss = multi.stats(), single.stats()
    assert sm["frames"] == 2 * F and ss["frames"] == F
    assert sm["iter_sum"] == 2 * ss["iter_sum"] and sm["iter_hist"] == [2 * v for v in ss["iter_hist"]]
    # device-pointer calls need ONE device
    with pytest.raises(q.QldpcError) as e:
        multi.decode_device(1, 0, 1, 1)
    assert e.value.code == 6
    multi.close()
    single.close()


def test_two_devices_through_the_c_abi(q, O, data_dir):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _case(q, O, data_dir, [0, 1])


def test_device_list_validation_and_single_entry_list(q, O, data_dir):
    code = q.Code.from_qc_file("%s/NR_1_1_384.qc" % data_dir)
    kw = dict(schedule=q.SCHED_LAYERED, rule=q.RULE_NMS, dtype=q.DTYPE_I8, max_iter=4, norm_factor=0.75)
    with pytest.raises(q.QldpcError):
        q.Decoder(code, devices=[0, 0], **kw)                  # a device twice
    with pytest.raises(q.QldpcError):
        q.Decoder(code, devices=[0, 99], **kw)
    d = q.Decoder(code, devices=[0], **kw)                     # a one-entry list is a plain single-device decoder
    llr = np.full((3, code.n), 20, np.int8)
    out, ok, iters, _ = d.decode(llr)
    assert ok.all() and not out.any()
    d.close()
