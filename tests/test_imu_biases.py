"""Static IMU biases (python/get_imu_biases.py of the reference), the producer of the hot CLI's --imu_bias_file.

Parity of this step is pinned by the REFERENCE ITSELF: tests/golden/imu_bias_*.npz hold inputs and the outputs of the reference's own
unmodified script (generator: tests/golden/make_imu_bias_golden.py).  CPU: the oracle reproduces them; GPU (-m gpu): the CUDA path
through the C-ABI and the drop-in tool reproduce them."""
import glob
import json
import os

import numpy as np
import pytest

from openimucameracalibrator_b200 import get_imu_biases as tool

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "imu_bias_*.npz")))
IDS = [os.path.basename(p) for p in GOLDEN]


def _inputs(z):
    a, g, t = tool.remove_seconds(z["accelerometer"], z["gyroscope"], z["timestamps_ns"], float(z["remove_sec"]))
    return np.asarray(a), np.asarray(g)


def test_fixtures_exist_and_trimming_matches_the_reference_importer():
    assert len(GOLDEN) >= 3
    z = np.load([p for p in GOLDEN if "trimmed" in p][0])
    a, g = _inputs(z)
    assert a.shape[0] == z["accelerometer"].shape[0] - 2 * round(float(z["remove_sec"]) * 400.0) and a.shape == g.shape


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_oracle_reproduces_reference_biases(oracle_factory, path):
    z = np.load(path)
    a, g = _inputs(z)
    ba, bg = oracle_factory().estimate_imu_biases(a, g, float(z["gravity_const"]))
    assert np.abs(ba - z["accl_bias"]).max() < 1e-13 and np.abs(bg - z["gyro_bias"]).max() < 1e-15


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_gpu_reproduces_reference_biases(gpu_factory, path):
    z = np.load(path)
    a, g = _inputs(z)
    ba, bg = gpu_factory().estimate_imu_biases(a, g, float(z["gravity_const"]))
    # tree reduction + atomics (order varies run to run) against NumPy's pairwise sums: means of ~10 m/s2 agree to a few 1e-14
    assert np.abs(ba - z["accl_bias"]).max() < 5e-13 and np.abs(bg - z["gyro_bias"]).max() < 1e-15


@pytest.mark.gpu
def test_gpu_bias_tool_writes_the_reference_file(tmp_path):
    z = np.load(GOLDEN[1])
    src, dst = str(tmp_path / "telemetry.json"), str(tmp_path / "imu_bias.json")
    json.dump({"accelerometer": z["accelerometer"].tolist(), "gyroscope": z["gyroscope"].tolist(), "timestamps_ns": z["timestamps_ns"].tolist(), "img_timestamps_ns": [],
               "camera_fps": 30.0}, open(src, "w"))
    tool.main(["--input_json_path", src, "--output_path", dst, "--gravity_const", repr(float(z["gravity_const"])), "--remove_sec", repr(float(z["remove_sec"]))])
    out = json.load(open(dst))
    assert set(out) == {"gyro_bias", "accl_bias"} and set(out["accl_bias"]) == {"x", "y", "z"}
    assert np.abs(np.array([out["accl_bias"][k] for k in "xyz"]) - z["accl_bias"]).max() < 5e-13
    assert np.abs(np.array([out["gyro_bias"][k] for k in "xyz"]) - z["gyro_bias"]).max() < 1e-15
