"""Host-side pieces of bench.py that need no GPU: the clock sampler's selection of the samples that fall inside the
timed region (the base contract wants clocks and throttle reasons sampled DURING it)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_clock_sampler_keeps_only_samples_of_the_timed_region():
    S = _bench().ClockSampler
    rows = [(0.10, 1200.0, 1965.0, set()), (0.50, 1965.0, 1965.0, {"sw_power_cap"}), (0.60, 1950.0, 1965.0, set()),
            (0.90, 600.0, 1965.0, {"hw_slowdown"})]
    got = S.summarise(rows, 0.45, 0.65, "nvml")
    assert got["samples"] == 2 and got["sm_mhz"] == 1957.5 and got["sm_max_mhz"] == 1965.0
    assert got["reasons"] == ["sw_power_cap"] and "note" not in got          # the slowdown outside the region is not reported
    near = S.summarise(rows, 0.62, 0.64, "nvml")                             # region shorter than the sampling period
    assert near["samples"] == 1 and near["sm_mhz"] == 1950.0 and "nearest sample" in near["note"]
    assert S.summarise([], 0.0, 1.0, "nvml")["reasons"] == ["no samples"]


def test_clock_sampler_without_nvml_or_nvidia_smi_reports_unavailable():
    s = _bench().ClockSampler(0)
    s.start(); s.mark_start(); s.mark_stop()
    out = s.stop()
    assert out["sm_mhz"] is None or out["samples"] >= 0                     # CPU box: unavailable; GPU box: real samples
