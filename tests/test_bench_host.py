"""Host-side logic of bench.py that needs no GPU: the step-level roofline bookkeeping and the clock sampler's windowing."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


class _Args:
    precision, bags, feat_size, dim, T = "bf16", 128, 1024, 512, 6


def test_step_roofline_matches_the_survey_figures():
    """cfg3 is 1 572 864 instance-passes x 4 591 360 FLOP = 7.22 TFLOP per step; the roof is the sum over the
    step's kernels of max(bytes / HBM peak, FLOP / tensor peak)."""
    peaks = {"hbm_gbs": 6540.5, "bf16_tflops_sustained": 1406.1}
    r = bench.step_roofline(_Args, 10.0, peaks)
    assert abs(r["flop"] - 1572864 * 4591360) / (1572864 * 4591360) < 0.01
    assert 38e9 < r["hbm_bytes"] < 41e9                                  # gather 4.8 GB + activations / gradients ~35 GB in bf16
    assert abs(r["hbm_only_ms"] - r["hbm_bytes"] / 6540.5e9 * 1e3) < 1e-2
    assert abs(r["tensor_only_ms"] - r["flop"] / 1406.1e12 * 1e3) < 1e-2
    assert max(r["hbm_only_ms"], r["tensor_only_ms"]) <= r["roof_ms"] <= r["hbm_only_ms"] + r["tensor_only_ms"]
    assert abs(r["frac"] - r["roof_ms"] / 10.0) < 1e-3
    # twice the slides: twice the work, same fraction at twice the time
    class Big(_Args):
        bags = 256
    r2 = bench.step_roofline(Big, 20.0, peaks)
    assert abs(r2["roof_ms"] - 2 * r["roof_ms"]) < 0.01 and abs(r2["frac"] - r["frac"]) < 1e-3
    json.dumps(r)                                                        # goes into the bench line as is


def test_clock_sampler_reports_only_the_timed_region():
    c = bench.Clocks(0)
    c.proc = type("P", (), {"terminate": lambda self: None})()          # no nvidia-smi here: feed the sampler's lines directly
    now = time.perf_counter()
    mk = lambda sm, cap: f"{sm}, 1965, Not Active, Not Active, Not Active, {cap}"
    c.lines = [(now - 1.0, mk(300, "Not Active")), (now - 0.5, mk(1965, "Not Active"))]      # before the region: idle clocks
    c.t_begin = now - 0.2
    c.lines += [(now - 0.15, mk(1700, "Active")), (now - 0.10, mk(1650, "Active")), (now - 0.05, mk(1680, "Active"))]
    out = c.stop()
    assert out["samples"] == 3 and out["samples_in_region"] == 3
    assert out["sm_mhz"] == 1680.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
    # a region shorter than the sampling period: the neighbours are used and the record says so
    c2 = bench.Clocks(0)
    c2.proc = c.proc
    now = time.perf_counter()
    c2.lines = [(now - 1.0, mk(1800, "Not Active"))]
    c2.t_begin = now - 0.001
    out2 = c2.stop()
    assert out2["samples_in_region"] == 0 and out2["samples"] == 1 and out2["sm_mhz"] == 1800.0
    assert bench.Clocks(0).stop()["reasons"] == ["nvidia-smi unavailable"]
