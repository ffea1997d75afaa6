"""bench.py host logic that can be checked without a GPU: the clock sampler reports only the nvidia-smi rows that fall inside a
timed window (it is started before the warm-up because nvidia-smi needs ~100 ms to deliver its first row)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def row(sm, mx=1965, power_cap="Not Active", thermal="Not Active"):
    return ["0", str(sm), str(mx), "700.0", "0x0", "Not Active", thermal, "Not Active", power_cap]


def test_clock_sampler_keeps_rows_inside_the_timed_windows_only():
    b = load_bench()
    s = b.ClockSampler(0)
    s.rows = [(0.5, row(300)), (1.10, row(1900)), (1.20, row(1950, power_cap="Active")), (1.9, row(500)), (2.55, row(1800)), (3.5, row(200))]
    s.windows = [(1.0, 1.5), (2.5, 2.6)]
    out = s.stop()
    assert out["samples"] == 3 and out["rows_total"] == 6
    assert out["sm_mhz"] == 1900 and out["sm_max_mhz"] == 1965
    assert out["reasons"] == ["sw_power_cap"]


def test_clock_sampler_without_rows_reports_none():
    b = load_bench()
    s = b.ClockSampler(0)
    s.windows = [(0.0, 1.0)]
    out = s.stop()
    assert out["samples"] == 0 and out["sm_mhz"] is None


def test_workload_name_states_the_sample_pipeline():
    b = load_bench()
    assert "simulate_modality + visual_perception_augmentation" in b.workload_name(True, True)
    assert "WITHOUT" in b.workload_name(False, False)
