"""Generates the committed golden fixtures under tests/golden/.

The reference (C++/Ceres/Theia) cannot be built or imported offline and ships no golden vectors, so
these fixtures are produced by the CPU ORACLE (oracle/) — they pin the oracle against silent drift and give the GPU tests
a reference that travels to the GPU box.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import F_ALL, F_STAGE1, F_STAGE2  # noqa: E402
from openimucameracalibrator_b200 import _capi as capi  # noqa: E402
from openimucameracalibrator_b200 import camera_models as cm  # noqa: E402
from openimucameracalibrator_b200 import synthetic as syn  # noqa: E402
from oracle_api import new_oracle  # noqa: E402

DS_KEYS = ["model", "intrinsics", "board_xyzw", "frame_t", "corner_offsets", "point_ids", "uv", "q_wc", "p_wc", "imu_t", "accel", "gyro", "dt_so3_s",
           "dt_r3_s", "std_so3", "std_r3", "time_offset_imu_to_cam_s", "init_line_delay_s", "T_i_c_init", "acc_bias", "gyr_bias", "gravity"]


def golden_cases():
    yield "tiny_division_undistortion", syn.tiny_config()
    yield "tiny_double_sphere_uneven_dt", syn.tiny_config(cm.DOUBLE_SPHERE, (342.4, 1.0, 0.0, 472.6, 273.9, -0.215, 0.513), dt_so3_s=0.04, dt_r3_s=0.07, seed=11)
    yield "tiny_extended_unified", syn.tiny_config(cm.EXTENDED_UNIFIED, (438.0, 1.0, 0.0, 489.5, 272.0, 0.5115, 1.062), n_frames=30, grid=(6, 5), seed=12)


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, cfg in golden_cases():
        ds = syn.make_dataset(cfg)
        rec = {k: np.asarray(ds[k]) for k in DS_KEYS}
        rec["image_size"] = np.asarray(ds["image_size"])
        o = new_oracle(1)
        capi.load_dataset(o, ds, known_gravity=False)
        for tag, flags in (("stage1", F_STAGE1), ("all", F_ALL), ("stage2", F_STAGE2)):
            c, r, g, H = o.evaluate(flags, hessian=True)
            rec[f"{tag}_cost"], rec[f"{tag}_residuals"], rec[f"{tag}_gradient"], rec[f"{tag}_hessian"] = c, r, g, H
        so3, r3, ba, bg = o.get_knots()
        rec["init_so3"], rec["init_r3"] = so3, r3
        o2 = new_oracle(1)
        capi.load_dataset(o2, ds, known_gravity=True)
        s1 = o2.optimize(50, F_STAGE1)
        rec["lm_stage1_iterations"], rec["lm_stage1_final_cost"], rec["lm_stage1_T_i_c"] = s1.iterations, s1.final_cost, o2.get_T_i_c()
        rec["lm_stage1_reproj"] = s1.mean_reproj_error
        s2 = o2.optimize(10, F_STAGE2)
        rec["lm_stage2_iterations"], rec["lm_stage2_line_delay"] = s2.iterations, o2.get_line_delay()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
        print(name, "frames", len(ds["frame_t"]), "cost", rec["stage1_cost"], "lm iters", s1.iterations, s2.iterations)


if __name__ == "__main__":
    main()
