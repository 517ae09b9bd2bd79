"""File formats + CLI boundary.  CPU: the Python writers round-trip and the C++ host program parses exactly what was written
(`--parse_only`, an extension flag) with the reference's flag syntax.  GPU: the full CLI run reproduces the direct C-ABI run."""
import json
import os
import subprocess

import numpy as np
import pytest

from helpers import F_STAGE1, F_STAGE2, rel
from openimucameracalibrator_b200 import _capi as capi
from openimucameracalibrator_b200 import camera_models as cm
from openimucameracalibrator_b200 import io_formats as iof
from openimucameracalibrator_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "openimucameracalibrator_b200", "bin", "continuous_time_imu_to_camera_calibration")


def _args(paths, out_dir, extra=()):
    a = [CLI]
    for k, v in paths.items():
        a.append(f"--{k}={v}")
    a += [f"--result_output_json={out_dir}/result.json", f"--output_path={out_dir}"]
    return a + list(extra)


def test_ubjson_roundtrip():
    doc = {"a": 1, "b": -300, "c": 70000, "d": 2 ** 40, "e": 1.5, "f": "text", "g": [1, 2.5, "x", [True, False, None]], "h": {"k": {"z": [0.1, 0.2]}}}
    assert iof.ubjson_loads(iof.ubjson_dumps(doc)) == doc


def test_view_key_matches_std_to_string():
    assert iof.view_key(1 / 30) == "33333.333333" and iof.view_key(0.0) == "0.000000"


@pytest.mark.parametrize("k", [0, 1, 2, 3, 4])
def test_cli_parses_written_files(tmp_path, k):
    assert os.path.exists(CLI), "build the CLI first (__graft_entry__.build())"
    cfg = syn.config5(k); cfg.n_frames = 8
    ds = syn.make_dataset(cfg)
    paths = iof.write_dataset_files(ds, str(tmp_path))
    out = subprocess.run(_args(paths, str(tmp_path), ["--parse_only", "--known_grav_dir_axis", "Z", "--nocalibrate_cam_line_delay"]), capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    s = json.loads(out.stdout)
    assert s["views"] == 8 and s["corners"] == 8 * 96 and s["imu_samples"] == len(ds["imu_t"]) and s["board_points"] == 96
    assert s["camera_model"] == ds["model"]
    k_expected = np.array(ds["intrinsics"], dtype=float)
    if ds["model"] not in (cm.DIVISION_UNDISTORTION, cm.FOV):
        k_expected[2] = 0.0                                   # the reference never reads `skew` (read_camera_calibration.cc)
    if ds["model"] == cm.PINHOLE:
        k_expected[5:] = 0.0                                  # ... nor PINHOLE radial distortion
    assert np.allclose(s["intrinsics"], k_expected)
    assert abs(s["uv_sum"] - float(np.sum(ds["uv"]))) < 1e-6
    assert abs(s["init_line_delay_s"] - 1.0 / ds["fps"] / ds["image_size"][1]) < 1e-18


def test_cli_reports_missing_inputs(tmp_path):
    out = subprocess.run([CLI, "--input_pose_dataset=/nonexistent.json", "--parse_only"], capture_output=True, text=True)
    assert out.returncode == 1 and "Could not read Reconstruction file" in out.stderr
    out = subprocess.run([CLI, "--no_such_flag=1"], capture_output=True, text=True)
    assert out.returncode == 1 and "unknown command line flag" in out.stderr


@pytest.mark.gpu
def test_cli_end_to_end_matches_direct_api(tmp_path, gpu_factory):
    cfg = syn.tiny_config(cm.DOUBLE_SPHERE, (342.4, 1.0, 0.0, 472.6, 273.9, -0.215, 0.513), n_frames=40, imu_rate_hz=200.0, seed=21, line_delay_init_scale=1.1)
    ds = syn.make_dataset(cfg)
    paths = iof.write_dataset_files(ds, str(tmp_path))
    out = subprocess.run(_args(paths, str(tmp_path), ["--calibrate_cam_line_delay", "--gravity_const=9.81"]), capture_output=True, text=True)
    assert out.returncode == 0, out.stderr + out.stdout
    res = iof.read_result_json(str(tmp_path / "result.json"))
    # the same problem through the C-ABI directly, with the values the files carry (timestamps rounded by the formats);
    # the CLI initialises the line delay to 1/fps/height itself (app :186), whatever the dataset's perturbed init was
    d2 = iof.dataset_from_files(ds)
    d2["init_line_delay_s"] = 1.0 / ds["fps"] / ds["image_size"][1]
    g = gpu_factory(); capi.load_dataset(g, d2)
    s1 = g.optimize(50, F_STAGE1); s2 = g.optimize(10, F_STAGE2)
    T = g.get_T_i_c()
    assert rel([res["q_i_c"]["x"], res["q_i_c"]["y"], res["q_i_c"]["z"], res["q_i_c"]["w"]], T[:4]) < 1e-9
    assert rel([res["t_i_c"]["x"], res["t_i_c"]["y"], res["t_i_c"]["z"]], T[4:]) < 1e-7
    assert abs(res["final_reproj_error"] - s1.mean_reproj_error) < 1e-9
    assert abs(res["calib_line_delay_us"] - g.get_line_delay() * 1e6) < 1e-9
    assert set(res) == {"q_i_c", "t_i_c", "final_reproj_error", "r3_dt", "so3_dt", "init_line_delay_us", "calib_line_delay_us", "time_offset_imu_to_cam_s", "trajectory"}
    t_used = g.imu_used()[0]
    assert len(res["trajectory"]) == len(np.unique((t_used * 1e9).astype(np.int64)))
    first = res["trajectory"][str(int(t_used[0] * 1e9))]
    assert set(first) == {"gyro_imu", "gyro_spline", "gyro_bias", "accl_imu", "accl_spline", "accl_bias"}
    assert os.path.exists(tmp_path / "sparse_recon_spline.ply") and os.path.exists(tmp_path / "sparse_recon_calib_dataset.ply")
