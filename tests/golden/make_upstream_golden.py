"""Generates the committed golden fixtures of the upstream rows f1 (board poses, board-point optimisation) and f4 (camera intrinsic
calibration) under tests/golden/upstream_*.npz.

Like the hot-path fixtures these are produced by the CPU ORACLE (the reference cannot be built offline and ships no vectors): they pin
the oracle against silent drift and give the GPU tests inputs + expected outputs that travel to the GPU box.
Run from the repo root:  python tests/golden/make_upstream_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_camera_calibration as tc  # noqa: E402
import test_pose_estimation as tp  # noqa: E402
from openimucameracalibrator_b200 import camera_models as cm  # noqa: E402
from oracle_api import new_oracle  # noqa: E402

TIGHT = dict(function_tolerance=1e-14, parameter_tolerance=1e-12, max_num_iterations=200)


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    rng = np.random.default_rng(2024)
    # ---- f1: poses of 40 noisy views (one gross outlier), then --optimize_board_points on a board known to 0.5 mm ------------------
    for name, (model, k) in (("fisheye", tp.CASES[1]), ("extended_unified", tp.CASES[4])):
        board, off, ids, uv, q_true, p_true = tp._scene(model, k, n_frames=40, seed=101 + model, noise_px=0.15)
        uv[off[3] + 17] += 45.0
        bent = board.copy(); bent[:, :3] += rng.normal(0, 5e-4, (board.shape[0], 3))
        o = new_oracle(1); o.set_camera(model, k, tp.W, tp.H); o.set_board_points(bent)
        q, p, e, v = o.estimate_board_poses(off, ids, uv)
        q2, p2, e2, v2, B2, n_opt = o.optimize_board_points(off, ids, uv, q, p, v)
        v3 = o.filter_bad_poses(p2, v2)
        np.savez_compressed(os.path.join(out_dir, f"upstream_f1_{name}.npz"), model=model, intrinsics=k, image_size=np.array([tp.W, tp.H]), board=bent, corner_offsets=off,
                            point_ids=ids, uv=uv, q_wc=q, p_wc=p, err=e, valid=v, q_wc_opt=q2, p_wc_opt=p2, err_opt=e2, valid_opt=v2, board_opt=B2, n_opt=n_opt, valid_filtered=v3)
        print("f1", name, "valid", int(v.sum()), "mean err", e[v > 0].mean(), "->", e2[v2 > 0].mean(), "optimised points", n_opt)
    # ---- f4: full calibration from the corners alone, with and without board-point refinement -------------------------------------
    for name, (model, k), opt in (("double_sphere", tc.CASES[3], 0), ("division_undistortion", tc.CASES[2], 0), ("extended_unified_boardopt", tc.CASES[4], 1)):
        B, off, ids, uv, q_true, p_true = tc.scene(model, k, n_views=24, seed=201 + model, noise_px=0.15)
        if opt:
            B = B.copy(); B[:, :3] += rng.normal(0, 3e-4, (B.shape[0], 3))
        uv[off[5]:off[6]] += rng.normal(0, 9.0, (B.shape[0], 2))        # one wrecked view: removed after stage 1
        o = new_oracle(1); o.set_board_points(B)
        r = o.calibrate_camera(model, tc.W, tc.H, off, ids, uv, grid_size=0.01, optimize_board_points=opt, **TIGHT)
        s = r["summary"]
        np.savez_compressed(os.path.join(out_dir, f"upstream_f4_{name}.npz"), model=model, image_size=np.array([tc.W, tc.H]), board=B, corner_offsets=off, point_ids=ids, uv=uv,
                            optimize_board_points=opt, grid_size=0.01, intrinsics=r["intrinsics"], q_wc=r["q_wc"], p_wc=r["p_wc"], view_error_px=r["view_error_px"], used=r["used"],
                            final_reproj_error=s["final_reproj_error"], final_cost=np.array(s["final_cost"]), focal_length_init=s["focal_length_init"],
                            n_views_selected=s["n_views_selected"], board_out=o.get_board_points(), truth=k)
        print("f4", name, "used", int(r["used"].sum()), "of", s["n_views_selected"], "f", r["intrinsics"][0], "truth", k[0], "err", s["final_reproj_error"])


if __name__ == "__main__":
    main()
