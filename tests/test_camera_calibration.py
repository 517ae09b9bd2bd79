"""Upstream row f4: camera intrinsic calibration, CameraCalibrator::CalibrateCameraFromJson / RunCalibration
(src/core/camera_calibrator.cc:131-389) = three theia::BundleAdjustViews passes over all view poses + one shared intrinsic vector.

CPU part: the oracle (Theia's formulation: angle-axis extrinsics, Jet autodiff, dense normal equations) is pinned by exact synthetic
truth -- noise-free projections of a planar board through the independent NumPy camera models are calibrated back to the generating
intrinsics and poses for every model whose staged schedule can reach them -- and by the reference's own rules (stage subsets,
view removal at 5 px / 2 px, the grid filter, the 10-view minimum).
GPU part (-m gpu): the CUDA path (closed-form Jacobians, quaternion right increments, per-view Schur elimination) lands on the
oracle's optimum on identical inputs: intrinsics to 1e-7 relative with tight tolerances, within the stopping tolerance otherwise."""
import time

import numpy as np
import pytest

from openimucameracalibrator_b200 import camera_models as cm
from openimucameracalibrator_b200 import synthetic as syn
from test_camera_models import CASES as MODEL_CASES

W, H = 960, 540
TIGHT = dict(function_tolerance=1e-14, parameter_tolerance=1e-12, max_num_iterations=200)


def _centred(model, k):
    """Same distortion, principal point at the image centre: the schedule of RunCalibration only frees the distortion in stage 1,
    where the principal point is pinned to the centre (:99, :146-160), so this is the case it can solve exactly."""
    k = np.array(k, dtype=np.float64)
    if model in (cm.FOV, cm.DIVISION_UNDISTORTION):
        k[2], k[3] = W / 2, H / 2
    else:
        k[3], k[4] = W / 2, H / 2
    return k


CASES = [(m, _centred(m, k)) for m, k in MODEL_CASES if m != cm.PINHOLE_RADIAL_TANGENTIAL]
IDS = [cm.MODEL_NAMES[m] for m, _ in CASES]


def _board(cols=9, rows=7, sq=0.021):
    gx, gy = np.meshgrid((np.arange(cols) - (cols - 1) / 2) * sq, (np.arange(rows) - (rows - 1) / 2) * sq)
    return np.stack([gx.ravel(), gy.ravel(), np.zeros(cols * rows), np.ones(cols * rows)], -1)


def scene(model, k, n_views=30, seed=0, noise_px=0.0, tilt=0.35, grid=(9, 7)):
    """Views with strong tilts and lateral offsets (focal length and distortion observable), all corners inside the image."""
    rng = np.random.default_rng(seed)
    B = _board(*grid)
    C = B.shape[0]
    q, p, uv = [], [], []
    while len(q) < n_views:
        R_wc = syn.so3_exp(np.array([np.pi, 0.0, 0.0]) + rng.normal(0, tilt, 3))
        pos = np.array([rng.uniform(-0.12, 0.12), rng.uniform(-0.07, 0.07), rng.uniform(0.22, 0.45)])
        px, valid = cm.project(model, k, (B[:, :3] - pos) @ R_wc)
        if not valid.all() or (px < 5).any() or (px[:, 0] > W - 5).any() or (px[:, 1] > H - 5).any():
            continue
        q.append(syn.matrix_to_quat_xyzw(R_wc[None])[0]); p.append(pos); uv.append(px + rng.normal(0, noise_px, px.shape))
    off = (np.arange(n_views + 1) * C).astype(np.int32)
    ids = np.tile(np.arange(C, dtype=np.int32), n_views)
    return B, off, ids, np.concatenate(uv), np.array(q), np.array(p)


def _qdiff(a, b):
    return np.minimum(np.abs(a - b).max(-1), np.abs(a + b).max(-1))


def _api(factory, board):
    a = factory(); a.set_board_points(board)
    return a


@pytest.mark.parametrize("model,k", CASES, ids=IDS)
def test_oracle_recovers_intrinsics_and_poses(oracle_factory, model, k):
    B, off, ids, uv, q_true, p_true = scene(model, k, seed=model)
    r = _api(oracle_factory, B).calibrate_camera(model, W, H, off, ids, uv, grid_size=0.0, **TIGHT)
    s = r["summary"]
    assert s["success"] == 1 and s["n_views_used"] == 30 and r["used"].all()
    assert np.abs(r["intrinsics"] - k).max() < 1e-6 * np.abs(k).max()
    assert _qdiff(r["q_wc"], q_true).max() < 1e-8 and np.abs(r["p_wc"] - p_true).max() < 1e-8
    assert s["final_reproj_error"] < 1e-8 and r["view_error_px"].max() < 1e-8


def test_oracle_stage_schedule_of_the_reference(oracle_factory):
    """Stage 3 frees principal point + focal length + aspect ratio, but the distortion only for PINHOLE (:184-194): with an
    off-centre principal point a DOUBLE_SPHERE calibration keeps the (biased) stage-1 xi / alpha, a PINHOLE one recovers everything."""
    model, k = cm.DOUBLE_SPHERE, np.array([342.4, 1.0, 0.0, 472.6, 273.9, -0.215, 0.513])
    B, off, ids, uv, q_true, p_true = scene(model, k, seed=11)
    o = _api(oracle_factory, B)
    full = o.calibrate_camera(model, W, H, off, ids, uv, grid_size=0.0, **TIGHT)
    one = o.calibrate_camera(model, W, H, off, ids, uv, grid_size=0.0, max_num_iterations=200, function_tolerance=1e-14, parameter_tolerance=1e-12,
                             max_view_error_stage1_px=1e9, max_view_error_final_px=1e9)
    assert full["summary"]["success"] == 1
    assert full["summary"]["final_cost"][2] <= full["summary"]["final_cost"][1] <= full["summary"]["final_cost"][0] * (1 + 1e-12)
    assert np.abs(full["intrinsics"][5:7] - k[5:7]).max() > 1e-4           # distortion frozen after stage 1: cannot be exact
    assert np.abs(full["intrinsics"][3:5] - k[3:5]).max() < 3.0            # ... the principal point still moves towards the truth
    assert np.allclose(one["intrinsics"], full["intrinsics"], rtol=1e-9)
    model, k = cm.PINHOLE, np.array([437.0, 1.02, 0.0, 489.0, 271.0, -0.05, 0.01])
    B, off, ids, uv, q_true, p_true = scene(model, k, seed=12)
    r = _api(oracle_factory, B).calibrate_camera(model, W, H, off, ids, uv, grid_size=0.0, **TIGHT)
    assert np.abs(r["intrinsics"] - k).max() < 1e-6 * 489.0


def test_oracle_noise_view_removal_grid_filter_and_minimum(oracle_factory):
    model, k = CASES[0]
    B, off, ids, uv, q_true, p_true = scene(model, k, n_views=24, seed=5, noise_px=0.2)
    C = B.shape[0]
    rng = np.random.default_rng(1)
    uv_bad = uv.copy()
    uv_bad[off[3]:off[4]] += rng.normal(0, 12.0, (C, 2))                     # a wrecked detection: mean error >> 5 px
    o = _api(oracle_factory, B)
    r = o.calibrate_camera(model, W, H, off, ids, uv_bad, grid_size=0.0)
    assert r["summary"]["success"] == 1 and r["used"][3] == 0 and r["used"].sum() == 23 and r["view_error_px"][3] == 0.0
    assert abs(r["intrinsics"][0] - k[0]) < 1.0 and np.abs(r["intrinsics"][3:5] - k[3:5]).max() < 1.5
    assert 0.15 < r["summary"]["final_reproj_error"] < 0.35                  # 0.2 px noise per axis -> mean norm ~0.25 px
    # grid filter (:313-325): repeating every view leaves the selection unchanged; a huge cell keeps a single view -> failure
    off2 = np.concatenate([off, off[1:] + off[-1]]).astype(np.int32)
    r2 = o.calibrate_camera(model, W, H, off2, np.tile(ids, 2), np.concatenate([uv, uv]), grid_size=1e-4)
    assert r2["summary"]["n_views_initialized"] == 48 and r2["summary"]["n_views_selected"] == 24 and not r2["used"][24:].any()
    r3 = o.calibrate_camera(model, W, H, off, ids, uv, grid_size=10.0)
    assert r3["summary"]["n_views_selected"] == 1 and r3["summary"]["success"] == 0
    # fewer than min_num_view_ = 10 views (camera_calibrator.h:84)
    r4 = o.calibrate_camera(model, W, H, off[:10], ids[: off[9]], uv[: off[9]], grid_size=0.0)
    assert r4["summary"]["success"] == 0 and r4["summary"]["n_views_selected"] == 9


def test_oracle_accepts_caller_initialisation(oracle_factory):
    """Poses + focal length handed in by the caller (the reference's AddView arguments, :84-95) replace the internal initialiser."""
    model, k = CASES[3]   # DOUBLE_SPHERE
    B, off, ids, uv, q_true, p_true = scene(model, k, seed=21)
    rng = np.random.default_rng(2)
    p_init = p_true + rng.normal(0, 2e-3, p_true.shape)
    r = _api(oracle_factory, B).calibrate_camera(model, W, H, off, ids, uv, q_wc_init=q_true, p_wc_init=p_init, focal_length_init=440.0, grid_size=0.0, **TIGHT)
    assert r["summary"]["init_iterations"] == 0 and r["summary"]["focal_length_init"] == 440.0
    assert np.abs(r["intrinsics"] - k).max() < 1e-6 * 480.0


# ------------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("model,k", CASES + [(cm.PINHOLE_RADIAL_TANGENTIAL, MODEL_CASES[-1][1])], ids=IDS + ["PINHOLE_RADIAL_TANGENTIAL"])
def test_gpu_calibration_matches_oracle(oracle_factory, gpu_factory, model, k):
    B, off, ids, uv, q_true, p_true = scene(model, k, n_views=36, seed=model + 40, noise_px=0.2)
    uv[off[5] + 7] += 25.0                                                   # a gross outlier inside the Huber tail
    ro = _api(oracle_factory, B).calibrate_camera(model, W, H, off, ids, uv, grid_size=0.0, **TIGHT)
    rg = _api(gpu_factory, B).calibrate_camera(model, W, H, off, ids, uv, grid_size=0.0, **TIGHT)
    so, sg = ro["summary"], rg["summary"]
    assert sg["success"] == 1 and so["success"] == 1 and (rg["used"] == ro["used"]).all() and sg["gpu_launches"] > 0
    assert abs(sg["focal_length_init"] - so["focal_length_init"]) < 1e-4 * so["focal_length_init"]
    n = cm.NUM_PARAMS[model]
    scale = np.maximum(np.abs(ro["intrinsics"]), 1e-3 * np.abs(ro["intrinsics"]).max())
    assert (np.abs(rg["intrinsics"] - ro["intrinsics"])[:n] / scale[:n]).max() < 1e-6
    assert _qdiff(rg["q_wc"], ro["q_wc"]).max() < 1e-7 and np.abs(rg["p_wc"] - ro["p_wc"]).max() < 1e-7
    assert np.abs(rg["view_error_px"] - ro["view_error_px"]).max() < 1e-6
    assert abs(sg["final_cost"][2] - so["final_cost"][2]) < 1e-7 * so["final_cost"][2]
    if model != cm.PINHOLE_RADIAL_TANGENTIAL:   # 0.2 px noise on 36 views; focal length and xi of the double sphere trade off
        assert abs(rg["intrinsics"][0] - k[0]) < (0.03 if model == cm.DOUBLE_SPHERE else 0.005) * k[0]


@pytest.mark.gpu
def test_gpu_calibration_default_tolerances_removal_and_init(oracle_factory, gpu_factory):
    model, k = CASES[4]   # EXTENDED_UNIFIED
    B, off, ids, uv, q_true, p_true = scene(model, k, n_views=30, seed=77, noise_px=0.2)
    rng = np.random.default_rng(3)
    uv[off[8]:off[9]] += rng.normal(0, 12.0, (B.shape[0], 2))
    ro = _api(oracle_factory, B).calibrate_camera(model, W, H, off, ids, uv)
    rg = _api(gpu_factory, B).calibrate_camera(model, W, H, off, ids, uv)
    assert (rg["used"] == ro["used"]).all() and rg["used"][8] == 0 and rg["summary"]["n_views_selected"] == ro["summary"]["n_views_selected"]
    # theia's default function tolerance 1e-6 stops both a little short of the optimum: agreement to the stopping tolerance
    assert np.abs(rg["intrinsics"] - ro["intrinsics"]).max() < 2e-3 * np.abs(ro["intrinsics"]).max()
    assert abs(rg["summary"]["final_reproj_error"] - ro["summary"]["final_reproj_error"]) < 1e-3
    # caller-provided initialisation
    p_init = p_true + rng.normal(0, 2e-3, p_true.shape)
    kw = dict(q_wc_init=q_true, p_wc_init=p_init, focal_length_init=430.0, grid_size=0.0, **TIGHT)
    ro = _api(oracle_factory, B).calibrate_camera(model, W, H, off, ids, uv, **kw)
    rg = _api(gpu_factory, B).calibrate_camera(model, W, H, off, ids, uv, **kw)
    assert rg["summary"]["init_iterations"] == 0 and (rg["used"] == ro["used"]).all()
    assert np.abs(rg["intrinsics"] - ro["intrinsics"]).max() < 1e-6 * np.abs(ro["intrinsics"]).max()
    # too few views: success = 0, status OK
    r = _api(gpu_factory, B).calibrate_camera(model, W, H, off[:8], ids[: off[7]], uv[: off[7]], grid_size=0.0)
    assert r["summary"]["success"] == 0
    g = gpu_factory()
    with pytest.raises(Exception):
        g.calibrate_camera(model, W, H, off, ids, uv)                        # board points not set


@pytest.mark.gpu
def test_gpu_calibration_full_size(oracle_factory, gpu_factory):
    """3000 views x 144 corners (BASELINE config 4 sizes, every view kept): converges to the generating intrinsics; wall time reported.
    A 40-view subset equals the oracle (the dense CPU formulation cannot hold 18 010 unknowns)."""
    model, k = CASES[4]
    B, off, ids, uv, q_true, p_true = scene(model, k, n_views=3000, seed=123, noise_px=0.2, grid=(16, 9))
    g = _api(gpu_factory, B)
    g.calibrate_camera(model, W, H, off[:41], ids[: off[40]], uv[: off[40]], grid_size=0.0)      # warm-up (allocations, module load)
    t0 = time.perf_counter()
    r = g.calibrate_camera(model, W, H, off, ids, uv, grid_size=0.0)
    dt = time.perf_counter() - t0
    s = r["summary"]
    assert s["success"] == 1 and s["n_views_used"] == 3000
    assert np.abs(r["intrinsics"] - k).max() < 2e-3 * np.abs(k).max()
    assert 0.2 < s["final_reproj_error"] < 0.3
    n = 40
    ro = _api(oracle_factory, B).calibrate_camera(model, W, H, off[: n + 1], ids[: off[n]], uv[: off[n]], grid_size=0.0, **TIGHT)
    rg = g.calibrate_camera(model, W, H, off[: n + 1], ids[: off[n]], uv[: off[n]], grid_size=0.0, **TIGHT)
    assert np.abs(rg["intrinsics"] - ro["intrinsics"]).max() < 1e-6 * np.abs(ro["intrinsics"]).max()
    its = sum(s["iterations"]) + s["init_iterations"]
    print(f"\n[f4] 3000 views x 144 corners: {dt * 1e3:.1f} ms wall incl. H2D/D2H, {its} LM iterations, {s['gpu_launches']} launches, "
          f"f = {r['intrinsics'][0]:.3f} (truth {k[0]}), mean reprojection error {s['final_reproj_error']:.4f} px")


@pytest.mark.gpu
def test_gpu_python_mirror_of_camera_calibrator(tmp_path):
    """openimucameracalibrator_b200.CameraCalibrator: the reference class's two entry routes give the C-ABI's numbers."""
    import json
    from openimucameracalibrator_b200 import CameraCalibrator
    model, k = CASES[1]   # FISHEYE
    B, off, ids, uv, q_true, p_true = scene(model, k, n_views=30, seed=31, noise_px=0.1)
    scene_json = {"camera_fps": 30.0, "image_width": W, "image_height": H, "scene_pts": {str(i): list(map(float, p[:3])) for i, p in enumerate(B)},
                  "views": {f"{f / 30.0 * 1e6:.6f}": {"image_points": {str(int(ids[c])): [float(uv[c, 0]), float(uv[c, 1])] for c in range(off[f], off[f + 1])}} for f in range(30)}}
    cal = CameraCalibrator("FISHEYE")
    cal.SetGridSize(0.001)
    assert cal.CalibrateCameraFromJson(scene_json, str(tmp_path / "cam"))
    doc = json.load(open(tmp_path / "cam.json"))
    assert doc["intrinsic_type"] == "FISHEYE" and doc["nr_calib_images"] == 30 and abs(doc["intrinsics"]["focal_length"] - k[0]) < 0.005 * k[0]
    assert abs(doc["intrinsics"]["radial_distortion_2"] - cal.result["intrinsics"][6]) < 1e-15
    cal.PrintResult()
    # AddView / AddObservation / RunCalibration with caller-made initial poses (R_cw, camera centre, focal length)
    cal2 = CameraCalibrator("FISHEYE"); cal2.SetBoardPoints(B)
    rng = np.random.default_rng(4)
    for f in range(30):
        R_cw = syn.quat_xyzw_to_matrix(q_true[f]).T
        vid = cal2.AddView(R_cw, p_true[f] + rng.normal(0, 1e-3, 3), 430.0, 0.0, W, H, f / 30.0)
        for c in range(off[f], off[f + 1]):
            assert cal2.AddObservation(vid, ids[c], uv[c])
    assert not cal2.AddObservation(99, 0, (0.0, 0.0))
    assert cal2.RunCalibration()
    assert np.abs(cal2.result["intrinsics"] - cal.result["intrinsics"]).max() < 2e-3 * k[0]     # default tolerances, two different starts


def _bent(board, sigma=3e-4, seed=0):
    bad = board.copy(); bad[:, :3] += np.random.default_rng(seed).normal(0, sigma, (board.shape[0], 3))
    return bad


def test_oracle_optimize_board_points_option(oracle_factory):
    """camera_calibrator.cc:207-216: BundleAdjustTracks on the board points + BundleAdjustViews again.  A board known to 0.3 mm leaves
    0.5 px of systematic error; refining the points removes it."""
    model, k = CASES[4]
    B, off, ids, uv, q_true, p_true = scene(model, k, n_views=30, seed=7)
    r0 = _api(oracle_factory, _bent(B)).calibrate_camera(model, W, H, off, ids, uv, grid_size=0.0, **TIGHT)
    o = _api(oracle_factory, _bent(B))
    r1 = o.calibrate_camera(model, W, H, off, ids, uv, grid_size=0.0, optimize_board_points=1, **TIGHT)
    assert r0["summary"]["n_points_optimized"] == 0 and r1["summary"]["n_points_optimized"] == B.shape[0]
    assert r0["summary"]["final_reproj_error"] > 0.3 and r1["summary"]["final_reproj_error"] < 0.1 * r0["summary"]["final_reproj_error"]
    assert not np.array_equal(o.get_board_points(), _bent(B))


@pytest.mark.gpu
def test_gpu_optimize_board_points_option_matches_oracle(oracle_factory, gpu_factory):
    model, k = CASES[4]
    B, off, ids, uv, q_true, p_true = scene(model, k, n_views=30, seed=8, noise_px=0.1)
    o, g = _api(oracle_factory, _bent(B)), _api(gpu_factory, _bent(B))
    ro = o.calibrate_camera(model, W, H, off, ids, uv, grid_size=0.0, optimize_board_points=1, **TIGHT)
    rg = g.calibrate_camera(model, W, H, off, ids, uv, grid_size=0.0, optimize_board_points=1, **TIGHT)
    assert rg["summary"]["n_points_optimized"] == ro["summary"]["n_points_optimized"] == B.shape[0] and (rg["used"] == ro["used"]).all()
    assert np.abs(g.get_board_points() - o.get_board_points()).max() < 1e-7
    assert np.abs(rg["intrinsics"] - ro["intrinsics"]).max() < 1e-6 * np.abs(ro["intrinsics"]).max()
    assert abs(rg["summary"]["final_reproj_error"] - ro["summary"]["final_reproj_error"]) < 1e-7
