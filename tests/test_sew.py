"""Upstream row f2: Spline Error Weighting, python/sew.py:knot_spacing_and_variance.

This is the one row whose parity is PINNED BY THE REFERENCE ITSELF: tests/golden/sew_*.npz hold inputs and outputs of the reference's
own python/sew.py (importable in the build container, see tests/golden/make_sew_golden.py).  The three cases cover the three exits of
find_max_quality_dt (python/sew.py:86-145): quality met at max_dt, Brent root inside the bracket, and "nothing satisfies it".
CPU: the oracle restatement (O(N^2) DFT) reproduces the fixtures.  GPU (-m gpu): the CUDA path (Bluestein FFT + fused reductions +
host Brent) reproduces the fixtures through the C-ABI."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "sew_*.npz")))
SETTINGS = (("r3", "accel", "q_r3", 0.01, 0.15), ("so3", "gyro", "q_so3", 0.01, 0.2))      # python/get_sew_for_dataset.py:38-48


def _check(api, path, rtol_dt, rtol_var):
    g = np.load(path)
    for tag, key, qk, lo, hi in SETTINGS:
        dt, var, spec = api.spline_error_weighting(g["times"], g[key], float(g[qk]), lo, hi, want_spectrum=True)
        assert abs(dt - float(g[f"{tag}_dt"])) <= rtol_dt * float(g[f"{tag}_dt"]), (tag, dt, float(g[f"{tag}_dt"]))
        assert abs(var - float(g[f"{tag}_var"])) <= rtol_var * float(g[f"{tag}_var"]), (tag, var, float(g[f"{tag}_var"]))
        head = g[f"{tag}_spectrum_head"]
        assert np.abs(spec[:16] - head).max() <= 1e-10 * np.abs(head).max()
        assert abs(np.sum(spec ** 2) / spec.size - float(g[f"{tag}_spectrum_energy"])) <= 1e-11 * float(g[f"{tag}_spectrum_energy"])


def test_fixtures_cover_all_three_exits():
    assert len(FILES) == 3
    dts = {os.path.basename(f): (float(np.load(f)["r3_dt"]), float(np.load(f)["so3_dt"])) for f in FILES}
    assert dts["sew_cfg2.npz"] == (0.15, 0.2)                               # quality already met at max_dt
    assert dts["sew_cfg3.npz"][0] == 0.01                                   # nothing satisfies it: best = min_dt
    assert 0.01 < dts["sew_bandlimited_prime.npz"][0] < 0.15 and 0.01 < dts["sew_bandlimited_prime.npz"][1] < 0.2    # Brent roots


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_the_reference(oracle_factory, path):
    _check(oracle_factory(), path, 1e-9, 1e-9)


def test_oracle_default_bounds(oracle_factory):
    g = np.load([f for f in FILES if "bandlimited" in f][0])
    dt, var = oracle_factory().spline_error_weighting(g["times"], g["gyro"], 0.98)          # min_dt = 1/rate, max_dt = n/4/rate
    assert abs(dt - float(g["so3_dt"])) < 1e-9 and abs(var - float(g["so3_var"])) < 1e-9 * float(g["so3_var"])    # same root from a wider bracket


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_gpu_reproduces_the_reference(gpu_factory, path):
    _check(gpu_factory(), path, 1e-9, 1e-9)


@pytest.mark.gpu
def test_gpu_sew_config4_size(oracle_factory, gpu_factory):
    """100 k samples (BASELINE config 4, 1 kHz): Bluestein length 262144; the first 20 k samples also go through the O(N^2) oracle."""
    import time
    from openimucameracalibrator_b200 import synthetic as syn
    ds = syn.make_dataset(syn.CONFIGS[4])
    t, acc, gyr = ds["imu_t"], ds["accel"].reshape(-1, 3), ds["gyro"].reshape(-1, 3)
    g = gpu_factory()
    g.spline_error_weighting(t, gyr, 0.98, 0.01, 0.2)
    t0 = time.perf_counter()
    r3 = g.spline_error_weighting(t, acc, 0.96, 0.01, 0.15); so3 = g.spline_error_weighting(t, gyr, 0.98, 0.01, 0.2)
    dt = time.perf_counter() - t0
    assert 0.01 <= r3[0] <= 0.15 and 0.01 <= so3[0] <= 0.2 and r3[1] > 0 and so3[1] > 0
    n = 20000
    og = oracle_factory().spline_error_weighting(t[:n], gyr[:n], 0.98, 0.01, 0.2); gg = g.spline_error_weighting(t[:n], gyr[:n], 0.98, 0.01, 0.2)
    assert abs(og[0] - gg[0]) <= 1e-9 * og[0] and abs(og[1] - gg[1]) <= 1e-9 * og[1]
    print(f"\n[f2] 2 x {t.size} samples (accelerometer + gyroscope): {dt * 1e3:.2f} ms wall incl. transfers; so3 dt {so3[0]:.4f} s, r3 dt {r3[0]:.4f} s")


@pytest.mark.gpu
def test_get_sew_for_dataset_tool(tmp_path):
    """The drop-in of python/get_sew_for_dataset.py: telemetry JSON in, spline_error_weighting JSON out (keys of :50-58)."""
    import json
    from openimucameracalibrator_b200 import get_sew_for_dataset as tool
    g = np.load([f for f in FILES if "bandlimited" in f][0])
    tel = {"accelerometer": g["accel"].tolist(), "gyroscope": g["gyro"].tolist(), "timestamps_ns": (g["times"] * 1e9).tolist(), "camera_fps": 0.0}
    src, dst = tmp_path / "telemetry.json", tmp_path / "sew.json"
    json.dump(tel, open(src, "w"))
    tool.main(["--input_json_path", str(src), "--output_path", str(dst)])
    out = json.load(open(dst))
    assert set(out) == {"so3", "r3", "camera_fps"} and set(out["so3"]) == {"knot_spacing", "weighting_factor", "quality_factor"}
    assert out["camera_fps"] == 30.0                                         # :55-58: fall-back when the telemetry has no fps
    # timestamps went through a float nanosecond round trip: 1e-6 relative is ample
    assert abs(out["so3"]["knot_spacing"] - float(g["so3_dt"])) < 1e-6 and abs(out["r3"]["weighting_factor"] - float(np.sqrt(g["r3_var"]))) < 1e-6
