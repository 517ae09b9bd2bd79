#!/usr/bin/env python
"""Workload for `ncu` / timing of row f4: camera intrinsic calibration of 3000 views x 144 corners (BASELINE config 4 sizes, every
view kept).  Prints the wall time of the second call (the first pays allocations) and the summary."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from openimucameracalibrator_b200 import _capi as capi, calibrator
from test_camera_calibration import CASES, W, H, scene
model, k = CASES[4]
B, off, ids, uv, q_true, p_true = scene(model, k, n_views=3000, seed=123, noise_px=0.2, grid=(16, 9))
g = capi.CApi(calibrator.load_library(), "icc_", 0); g.set_board_points(B)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for rep in range(reps):
    t0 = time.perf_counter(); r = g.calibrate_camera(model, W, H, off, ids, uv, grid_size=0.0); dt = time.perf_counter() - t0
s = r["summary"]
print(f"f4: {dt * 1e3:.2f} ms wall, iterations {s['iterations']} + init {s['init_iterations']}, launches {s['gpu_launches']}, f = {r['intrinsics'][0]:.4f}, reproj {s['final_reproj_error']:.4f} px")
