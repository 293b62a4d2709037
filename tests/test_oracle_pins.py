"""The CPU oracle against the results the reference RECORDED for its own decoders (SURVEY.md 8c pin 4, tests/golden/
recorded_fer.json; workloads in tests/refpins.py).  No GPU.  Frame counts are cut so that the file runs in under a minute;
tests/test_gpu_pins.py repeats every point at full size on the GPU, bit-exact with this oracle."""
import numpy as np
import pytest

import refpins


def _oracle_nr_encoder(oc):
    return lambda msgs: np.stack([oc.nr_encode(m) for m in msgs])


@pytest.mark.parametrize("name,idx,frames", [("NR_1_1_24", 0, 1500), ("NR_1_1_24", 1, 3000), ("NR_1_1_24", 2, 12000),
                                             ("NR_2_6_52", 0, 1500), ("NR_2_6_52", 1, 3000), ("NR_2_6_52", 2, 8000)])
def test_matlab_rm_fp_recorded_fer(O, data_dir, name, idx, frames):
    """integer layered offset min-sum (ML/BPSK_nrldpc_sim_RM_FP.m) against the frame-error counts in ML/sim_results.m:2-5,
    9-12, with the rate definition k/n (see refpins.py).  Both sides are samples (the recorded one of 100..5000 frames, this
    one of `frames`), so the criterion per point is that the two Clopper-Pearson 95 % intervals overlap; the GPU test runs
    the full table at 20 000 frames per point and also counts the estimates that fall inside the recorded interval."""
    row = refpins.recorded()["sim_results_m"][name][idx]
    full = O.Code.from_qc("%s/%s.qc" % (data_dir, name))
    kb, nbRM, mbRM = refpins.rm_fp_geometry(full.base)
    sub = O.Code.from_base(full.base[:mbRM, :nbRM].copy(), full.Z)
    msgs, llr = refpins.rm_fp_frames(full.base, full.Z, row["ebno_db"], frames, _oracle_nr_encoder(full), seed=1)
    hard, it, ok, _ = sub.batch_layered_fixed_i8(llr, None, rule=O.RULE_OMS, n_ite=20, early_stop=False, offset=2,
                                                 msg_max=31, app_max=127)
    assert (it == 20).all()
    fe = int((hard[:, :sub.K] != msgs).any(axis=1).sum())
    assert refpins.consistent(fe, frames, row["frame_errors"], row["frames"]), \
        "FER %.4f (%d / %d) vs recorded %.4f (%d / %d)" % (fe / frames, fe, frames, row["fer"], row["frame_errors"], row["frames"])


def test_matlab_rm_fp_committed_rate_definition_is_not_what_was_recorded(O, data_dir):
    """the script as committed (sigma from k/(n-2z), :24) decodes clearly BETTER than the recorded table: the table was made
    with k/n.  Kept as a test so that the documented discrepancy stays true."""
    row = refpins.recorded()["sim_results_m"]["NR_2_6_52"][1]            # 1.5 dB: recorded 335 / 1000
    full = O.Code.from_qc("%s/NR_2_6_52.qc" % data_dir)
    kb, nbRM, mbRM = refpins.rm_fp_geometry(full.base)
    sub = O.Code.from_base(full.base[:mbRM, :nbRM].copy(), full.Z)
    msgs, llr = refpins.rm_fp_frames(full.base, full.Z, 1.5, 2000, _oracle_nr_encoder(full), seed=1, rate_def="punctured")
    hard, _, _, _ = sub.batch_layered_fixed_i8(llr, None, rule=O.RULE_OMS, n_ite=20, early_stop=False, offset=2)
    fer = float((hard[:, :sub.K] != msgs).any(axis=1).mean())
    lo, _ = refpins.clopper_pearson(row["frame_errors"], row["frames"])
    assert fer < 0.5 * lo


def test_5g_qc_driver_recorded_fer_and_ber(O, data_dir):
    """float flooding SPA, the (5g-qc) driver's set-up on NR_1_0_2.qc: every one of the 11 recorded QBER steps
    (README_LDPC.md:941-974, 100 frames each) inside its 95 % interval, recorded puncture counts reproduced, and the
    information-bit error rate within 20 % of the recorded one summed over the sweep (BER separates SPA from min-sum:
    the min-sum rules give 2-6 x the recorded BER at the same FER)."""
    rows = [r for r in refpins.recorded()["readme_ldpc_nr_1_0_2"]["rows"] if r["qber"] > 0]
    assert len(rows) == 11
    oc = O.Code.from_qc("%s/NR_1_0_2.qc" % data_dir)
    assert (oc.N, oc.K) == (136, 44)
    F = 1500
    inside = 0
    be_sum = be_ref = 0.0
    for row in rows:
        msgs, llr, punct = refpins.qc5g_frames(oc.N, oc.K, row["qber"], F, _oracle_nr_encoder(oc), seed=int(row["qber"] * 100))
        assert punct == row["punctured_bits"]
        hard, post, it, ok, _ = oc.batch_flooding_f32(llr, None, rule=O.RULE_SPA, n_ite=10, early_stop=True)
        fe = int((hard[:, :oc.K] != msgs).any(axis=1).sum())
        lo, hi = refpins.clopper_pearson(row["frame_errors"], row["frames"])
        inside += lo <= fe / F <= hi
        be_sum += float((hard[:, :oc.K] != msgs).mean())
        be_ref += row["ber"]
    assert inside == 11, "%d of 11 QBER steps inside their interval" % inside
    assert abs(be_sum / be_ref - 1) < 0.2
