"""The measurement / rate-adaptation tools under tools/ run end to end on small inputs and report sane numbers: the
puncture-pattern search of the reference driver (BOOT/src/main.cpp:305-409) as a GPU-batched tool, the config 3 probes, the
layered probe, the config 4 stream."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=300):
    p = subprocess.run([sys.executable] + args, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stderr[-2000:]
    return [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]


def test_puncture_search_finds_a_clean_pattern_and_counts_leakage():
    rows = _run(["tools/puncture_search.py", "--frames", "256", "--patterns", "3", "--puncture", "3000", "--qber", "0.02"])
    summary = rows[-1]
    assert summary["patterns_tried"] >= 1 and summary["clean_pattern"] is not None
    n, k = 26112, 8448
    assert abs(summary["leak_per_key_bit"] - (n - k - 3000) / k) < 1e-9 and summary["efficiency_f"] > 1.0
    assert all(r["frames"] == 256 for r in rows[:-1])
    # a pattern that punctures nearly every parity bit cannot be clean at this QBER: the search reports that, too
    rows = _run(["tools/puncture_search.py", "--frames", "64", "--patterns", "2", "--puncture", "16500", "--qber", "0.05"])
    assert rows[-1]["clean_pattern"] is None and rows[-1]["patterns_tried"] == 2 and rows[0]["frame_errors"] > 0


def test_flooding_probe_reconciles_on_both_long_codes():
    for code, frames in (("qkd_psdpeg_n65536.qc", "16"), ("qkd_irregular_n65536_r34.qc", "8")):
        rows = _run(["tools/flood_bench.py", "--code", code, "--frames", frames, "--reps", "1", "--warmup", "1", "--qber", "0.02"])
        assert len(rows) == 5 and all(r["kernel"] == "flooding_qc_cluster" for r in rows)
        assert all(r["all_reconciled"] and r["ok_frac"] == 1.0 for r in rows), rows
        exact, fast = rows[0], rows[1]
        assert exact["mean_sweeps"] == fast["mean_sweeps"]            # fp32 SPA: same iteration counts as the exact flavour


def test_layered_probe_covers_every_layered_kernel():
    rows = _run(["tools/layered_bench.py", "--frames", "64", "--reps", "1", "--iters", "6"])
    kernels = {r["kernel"] for r in rows}
    assert {"layered_i8s_zpack4", "layered_i8_zpack4", "layered_generic", "layered_csr"} <= kernels
    assert all(r["iterations"] == 6.0 and r["edge_updates_per_s"] > 0 for r in rows)
    assert all(r["all_reconciled"] for r in rows if "f32" in r["case"] or "i16" in r["case"])


def test_layered_probe_with_early_termination():
    """the usual mode of operation: syndrome test after every iteration (bit-vector test in layered_generic for Z % 32 == 0,
    per-lane test otherwise); every float / int16 frame still reconciles, in fewer iterations than the fixed count"""
    rows = _run(["tools/layered_bench.py", "--frames", "128", "--reps", "1", "--iters", "10", "--early-stop", "--only", "Z="])
    gen = [r for r in rows if r["kernel"] == "layered_generic"]
    assert len(gen) >= 4
    assert all(r["all_reconciled"] for r in gen if "f32" in r["case"] or "i16" in r["case"])
    assert all(1.0 <= r["iterations"] < 6.0 for r in gen), [(r["case"], r["iterations"]) for r in gen]


def test_stream_tool_counts_every_frame():
    rows = _run(["tools/stream_10gbit.py", "--gbit", "0.2", "--chunk", "8192"])
    r = rows[-1]
    frames = r["config"]["frames"]
    assert frames == -(-200000000 // 8448) and r["decoder_stats"]["frames"] == frames == r["decoder_stats"]["frames_expected"]
    assert r["frames_not_reconciled"] == 0 and r["fer"] == 0.0 and 1.0 < r["decoder_stats"]["mean_iters"] < 3.0 and r["value"] > 0
