"""Upstream row f1: per-view board poses, PoseEstimator::EstimatePosesFromJson (src/core/pose_estimator.cc:92-191).

CPU part: the oracle restatement is pinned by exact synthetic truth (noise-free global-shutter projections of a planar board through
the independent NumPy camera models): un-projection is the inverse of projection for all seven models, the estimated pose is the
generating pose, the reference's drop rules (min 8 corners, >= 6 inliers) and the outlier threshold behave as written.
GPU part (-m gpu): the CUDA path through the C-ABI reproduces the oracle on identical inputs, bit-for-bit in validity and to 2e-8
in pose (1e-9 on noise-free views); the converged pose does not depend on the start (both sides solve the same convex-near-the-optimum problem to 1e-15)."""
import numpy as np
import pytest

from openimucameracalibrator_b200 import _capi as capi
from openimucameracalibrator_b200 import camera_models as cm
from openimucameracalibrator_b200 import synthetic as syn
from test_camera_models import CASES

W, H = 960, 540


def _board(cols=9, rows=7, sq=0.021, z=0.0):
    gx, gy = np.meshgrid((np.arange(cols) - (cols - 1) / 2) * sq, (np.arange(rows) - (rows - 1) / 2) * sq)
    return np.stack([gx.ravel(), gy.ravel(), np.full(cols * rows, z), np.ones(cols * rows)], -1)


def _scene(model, k, n_frames=12, seed=0, noise_px=0.0, z_plane=0.0):
    """Random camera poses ~0.35-0.5 m above a planar board; exact (global-shutter) projections."""
    rng = np.random.default_rng(seed)
    board = _board(z=z_plane)
    C = board.shape[0]
    q_wc, p_wc, uv = [], [], []
    while len(q_wc) < n_frames:
        R_wc = syn.so3_exp(np.array([np.pi, 0.0, 0.0]) + rng.normal(0, 0.15, 3))       # camera looks down (-z world)
        p = np.array([rng.uniform(-0.03, 0.03), rng.uniform(-0.03, 0.03), z_plane + rng.uniform(0.35, 0.5)])
        pc = (board[:, :3] - p) @ R_wc                                                   # R_wc^T (X - p)
        px, valid = cm.project(model, k, pc)
        if not valid.all() or (px < 0).any() or (px[:, 0] > W).any() or (px[:, 1] > H).any():
            continue
        q_wc.append(syn.matrix_to_quat_xyzw(R_wc[None])[0]); p_wc.append(p); uv.append(px + rng.normal(0, noise_px, px.shape))
    off = (np.arange(n_frames + 1) * C).astype(np.int32)
    ids = np.tile(np.arange(C, dtype=np.int32), n_frames)
    return board, off, ids, np.concatenate(uv), np.array(q_wc), np.array(p_wc)


def _setup(api, model, k, board):
    api.set_camera(model, k, W, H); api.set_board_points(board)
    return api


def _qdiff(a, b):
    return np.minimum(np.abs(a - b).max(-1), np.abs(a + b).max(-1))


@pytest.mark.parametrize("model,k", CASES, ids=[cm.MODEL_NAMES[m] for m, _ in CASES])
def test_oracle_unprojection_inverts_projection(oracle_factory, model, k):
    rng = np.random.default_rng(model + 10)
    pts = np.concatenate([rng.uniform(-0.45, 0.45, size=(300, 2)), np.ones((300, 1))], axis=1)
    px, valid = cm.project(model, k, pts)
    o = _setup(oracle_factory(), model, k, _board())
    xy, ok = o.pixels_to_normalized(px[valid])
    assert ok.all()
    assert np.abs(xy - pts[valid, :2]).max() < 1e-10


@pytest.mark.parametrize("model,k", CASES, ids=[cm.MODEL_NAMES[m] for m, _ in CASES])
def test_oracle_recovers_exact_poses(oracle_factory, model, k):
    board, off, ids, uv, q_true, p_true = _scene(model, k, seed=model)
    q, p, e, v = _setup(oracle_factory(), model, k, board).estimate_board_poses(off, ids, uv)
    assert v.all()
    assert _qdiff(q, q_true).max() < 1e-9 and np.abs(p - p_true).max() < 1e-9 and e.max() < 1e-10


def test_oracle_plane_offset_and_noise(oracle_factory):
    model, k = CASES[4]
    board, off, ids, uv, q_true, p_true = _scene(model, k, seed=3, noise_px=0.2, z_plane=0.25)
    q, p, e, v = _setup(oracle_factory(), model, k, board).estimate_board_poses(off, ids, uv)
    assert v.all()
    assert _qdiff(q, q_true).max() < 1e-2 and np.abs(p - p_true).max() < 1e-2     # 0.2 px on a 17 cm board at 0.4 m: tilt is weakly observable
    assert 2e-4 < e.mean() < 1.2e-3                                                # ~0.25 px / 438 px focal, normalised units


def test_oracle_drop_rules_and_outliers(oracle_factory):
    model, k = CASES[0]
    board, off, ids, uv, q_true, p_true = _scene(model, k, n_frames=4, seed=5)
    C = board.shape[0]
    # view 1: only 7 corners (min_num_points_ = 8, pose_estimator.h:72) ; view 2: 8 corners is enough
    keep = np.ones(uv.shape[0], bool); keep[C + 7:2 * C] = False
    keep[2 * C:3 * C] = False; keep[2 * C + np.array([0, 4, 8, 13, 22, 31, 44, 62])] = True   # 8 corners spread over the board
    off2 = np.array([0, C, C + 7, C + 15, 2 * C + 15], np.int32)
    uv2, ids2 = uv[keep].copy(), ids[keep]
    # view 3: two gross outliers (50 px) must be rejected by the squared normalised threshold and not bias the pose
    uv2[off2[3] + 3] += 50.0; uv2[off2[3] + 40] -= 50.0
    q, p, e, v = _setup(oracle_factory(), model, k, board).estimate_board_poses(off2, ids2, uv2)
    assert v.tolist() == [1, 0, 1, 1]
    assert np.abs(p[[0, 2, 3]] - p_true[[0, 2, 3]]).max() < 1e-8 and _qdiff(q[[0, 2, 3]], q_true[[0, 2, 3]]).max() < 1e-8
    assert (q[1] == [0, 0, 0, 1]).all() and (p[1] == 0).all()
    # a non-planar target is refused (the initialisation is the planar homography)
    bad = board.copy(); bad[5, 2] = 0.01
    q, p, e, v = _setup(oracle_factory(), model, k, bad).estimate_board_poses(off, ids, uv)
    assert not v.any()


def test_estimated_poses_feed_the_calibration(oracle_factory):
    """End to end on the CPU oracle: poses estimated from the corners replace the dataset's pose priors and the LM solve lands on
    the same calibration (the priors only seed the spline knots, impl.h:278-339)."""
    from helpers import F_STAGE1
    ds = syn.make_dataset(syn.tiny_config())
    o = oracle_factory(); capi.load_dataset(o, ds)
    ref = o.optimize(30, F_STAGE1); T_ref = o.get_T_i_c()
    o2 = oracle_factory()
    o2.set_camera(ds["model"], ds["intrinsics"], *ds["image_size"]); o2.set_board_points(ds["board_xyzw"])
    q, p, e, v = o2.estimate_board_poses(ds["corner_offsets"], ds["point_ids"], ds["uv"])
    assert v.all()
    ds2 = dict(ds, q_wc=q, p_wc=p)
    o3 = oracle_factory(); capi.load_dataset(o3, ds2)
    s = o3.optimize(30, F_STAGE1)
    # Ceres-style stop at function_tolerance 1e-4: two starts end within a few per cent of each other, not at the same digits
    assert s.termination in (1, 2, 3) and abs(s.final_cost - ref.final_cost) < 3e-2 * ref.final_cost
    assert np.abs(o3.get_T_i_c() - T_ref).max() < 5e-3


# ------------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("model,k", CASES, ids=[cm.MODEL_NAMES[m] for m, _ in CASES])
def test_gpu_unprojection_matches_oracle(oracle_factory, gpu_factory, model, k):
    rng = np.random.default_rng(model + 20)
    px = np.stack([rng.uniform(0, W, 500), rng.uniform(0, H, 500)], -1)
    xo, oko = _setup(oracle_factory(), model, k, _board()).pixels_to_normalized(px)
    xg, okg = _setup(gpu_factory(), model, k, _board()).pixels_to_normalized(px)
    assert (oko == okg).all() and oko.sum() > 400
    m = oko.astype(bool)
    assert np.abs(xg[m] - xo[m]).max() < 1e-10 * max(1.0, np.abs(xo[m]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("model,k", CASES, ids=[cm.MODEL_NAMES[m] for m, _ in CASES])
def test_gpu_poses_match_oracle(oracle_factory, gpu_factory, model, k):
    board, off, ids, uv, q_true, p_true = _scene(model, k, n_frames=40, seed=model + 30, noise_px=0.2)
    uv[off[7] + 11] += 60.0                                     # one gross outlier in view 7
    qo, po, eo, vo = _setup(oracle_factory(), model, k, board).estimate_board_poses(off, ids, uv)
    qg, pg, eg, vg = _setup(gpu_factory(), model, k, board).estimate_board_poses(off, ids, uv)
    assert (vo == vg).all() and vo.all()
    # both sides stop when the cost stalls at 1e-15 relative, which pins the pose of a noisy view to ~1e-9
    assert _qdiff(qg, qo).max() < 2e-8 and np.abs(pg - po).max() < 2e-8 and np.abs(eg - eo).max() < 1e-10
    assert _qdiff(qg, q_true).max() < 1.5e-2 and np.abs(pg - p_true).max() < 1.5e-2


@pytest.mark.gpu
def test_gpu_pose_drop_rules(gpu_factory):
    model, k = CASES[0]
    board, off, ids, uv, q_true, p_true = _scene(model, k, n_frames=3, seed=9)
    C = board.shape[0]
    keep = np.ones(uv.shape[0], bool); keep[C + 7:2 * C] = False
    off2 = np.array([0, C, C + 7, 2 * C + 7], np.int32)
    q, p, e, v = _setup(gpu_factory(), model, k, board).estimate_board_poses(off2, ids[keep], uv[keep])
    assert v.tolist() == [1, 0, 1] and (q[1] == [0, 0, 0, 1]).all()
    assert np.abs(p[[0, 2]] - p_true[[0, 2]]).max() < 1e-8
    g = gpu_factory()
    with pytest.raises(Exception):
        g.estimate_board_poses(off, ids, uv)                   # camera / board not set


@pytest.mark.gpu
def test_gpu_poses_config4_full_size(oracle_factory, gpu_factory):
    """3000 views x 144 corners (BASELINE config 4): all views valid, a 60-view sample equals the oracle, device time reported."""
    import time
    ds = syn.make_dataset(syn.CONFIGS[4])
    g = gpu_factory()
    g.set_camera(ds["model"], ds["intrinsics"], *ds["image_size"]); g.set_board_points(ds["board_xyzw"])
    g.estimate_board_poses(ds["corner_offsets"], ds["point_ids"], ds["uv"])             # warm-up (allocations)
    t0 = time.perf_counter()
    q, p, e, v = g.estimate_board_poses(ds["corner_offsets"], ds["point_ids"], ds["uv"])
    dt = time.perf_counter() - t0
    assert v.all() and np.isfinite(q).all()
    # rolling-shutter corners against a global-shutter pose: centimetre-level agreement with the (noisy) priors of the generator
    assert np.abs(p - ds["p_wc"]).max() < 0.05 and _qdiff(q, ds["q_wc"]).max() < 0.05
    n = 60
    off = ds["corner_offsets"][: n + 1]
    o = oracle_factory()
    o.set_camera(ds["model"], ds["intrinsics"], *ds["image_size"]); o.set_board_points(ds["board_xyzw"])
    qo, po, eo, vo = o.estimate_board_poses(off, ds["point_ids"][: off[-1]], ds["uv"][: off[-1]])
    assert _qdiff(q[:n], qo).max() < 2e-8 and np.abs(p[:n] - po).max() < 2e-8
    print(f"\n[f1] 3000 views x 144 corners: {dt * 1e3:.2f} ms wall incl. H2D/D2H ({3000 / dt:.0f} views/s)")


# ---- --optimize_board_points of the pose app + FilterBadPoses (app :61-67) ----------------------------------------------------------
def _perturbed(board, sigma=5e-4, seed=0):
    bad = board.copy(); bad[:, :3] += np.random.default_rng(seed).normal(0, sigma, (board.shape[0], 3))
    return bad


def test_oracle_board_point_optimisation(oracle_factory):
    """OptimizeBoardPoints + OptimizeAllPoses (pose_estimator.cc:193-236): exact data is a fixed point; a board known only to 0.5 mm
    is pulled towards consistency with the images (the poses were estimated on the wrong board, so the truth itself is out of reach)."""
    model, k = CASES[4]
    board, off, ids, uv, q_true, p_true = _scene(model, k, n_frames=60, seed=3)
    o = _setup(oracle_factory(), model, k, board)
    q, p, e, v, B, n = o.optimize_board_points(off, ids, uv, q_true, p_true, np.ones(60, np.int32))
    assert n == board.shape[0] and v.all()
    assert np.abs(B[:, :3] - board[:, :3]).max() < 1e-12 and np.abs(p - p_true).max() < 1e-10 and e.max() < 1e-10
    assert np.array_equal(o.get_board_points(), B)
    # only points seen in more than min_num_obs_for_optim_ = 30 views are touched (pose_estimator.h:78, .cc:202-206)
    keep = np.ones(uv.shape[0], bool); C = board.shape[0]
    for f in range(35, 60):
        keep[f * C + 5] = False                                       # point 5 is seen in 35 views, point 6 in 25
    for f in range(25, 60):
        keep[f * C + 6] = False
    off2 = np.concatenate([[0], np.cumsum([keep[f * C:(f + 1) * C].sum() for f in range(60)])]).astype(np.int32)
    o = _setup(oracle_factory(), model, k, _perturbed(board))
    q0, p0, e0, v0 = o.estimate_board_poses(off2, ids[keep], uv[keep])
    before = o.get_board_points()
    q, p, e, v, B, n = o.optimize_board_points(off2, ids[keep], uv[keep], q0, p0, v0)
    assert n == C - 1 and np.array_equal(B[6], before[6]) and not np.array_equal(B[5], before[5])
    assert v0.all() and v.all() and e[v > 0].mean() < 0.1 * e0[v0 > 0].mean()


def test_filter_bad_poses(oracle_factory):
    """PoseEstimator::FilterBadPoses (pose_estimator.cc:238-261) -- host logic, identical in the product library and the oracle."""
    from openimucameracalibrator_b200 import calibrator
    p = np.array([[0, 0, 0.40], [0, 0, 0.45], [0, 0, -0.40], [0, 0, 0.90], [0, 0, 0.41], [5, 5, 0.0]])
    valid = [1, 1, 1, 1, 1, 0]
    expect = [1, 1, 0, 0, 1, 0]                                       # median of the valid heights 0.41: |z - 0.41| > 0.41 drops -0.40 and 0.90
    assert oracle_factory().filter_bad_poses(p, valid).tolist() == expect
    h = capi.CApi(calibrator.load_library(), "icc_", -1)
    assert h.filter_bad_poses(p, valid).tolist() == expect
    assert h.filter_bad_poses(p[:2], [0, 0]).tolist() == [0, 0]


@pytest.mark.gpu
def test_gpu_board_point_optimisation_matches_oracle(oracle_factory, gpu_factory):
    model, k = CASES[4]
    board, off, ids, uv, q_true, p_true = _scene(model, k, n_frames=60, seed=4, noise_px=0.1)
    uv[off[9] + 20] += 40.0
    bad = _perturbed(board, seed=1)
    o, g = _setup(oracle_factory(), model, k, bad), _setup(gpu_factory(), model, k, bad)
    qo, po, eo, vo = o.estimate_board_poses(off, ids, uv)
    qg, pg, eg, vg = g.estimate_board_poses(off, ids, uv)
    ro = o.optimize_board_points(off, ids, uv, qo, po, vo)
    rg = g.optimize_board_points(off, ids, uv, qg, pg, vg)
    assert rg[5] == ro[5] == board.shape[0] and (rg[3] == ro[3]).all() and ro[3].all()
    assert np.abs(rg[4] - ro[4]).max() < 1e-8 and np.abs(rg[1] - ro[1]).max() < 1e-7 and _qdiff(rg[0], ro[0]).max() < 1e-7
    assert np.abs(rg[2] - ro[2]).max() < 1e-9 and rg[2].mean() < 0.5 * eg.mean()
    assert np.array_equal(g.get_board_points(), rg[4])
    # exact data: a fixed point of the GPU path too; refine-only poses == a fresh estimate
    board, off, ids, uv, q_true, p_true = _scene(model, k, n_frames=40, seed=6)
    g = _setup(gpu_factory(), model, k, board)
    q, p, e, v, B, n = g.optimize_board_points(off, ids, uv, q_true, p_true, np.ones(40, np.int32))
    assert n == board.shape[0] and np.abs(B[:, :3] - board[:, :3]).max() < 1e-10 and np.abs(p - p_true).max() < 1e-9
