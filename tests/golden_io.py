import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "tiny_*.npz")))     # (sew_*.npz belong to tests/test_sew.py)


def load_golden(path):
    z = np.load(path)
    ds = {k: z[k] for k in z.files}
    for k in ("model",):
        ds[k] = int(ds[k])
    for k in ("dt_so3_s", "dt_r3_s", "std_so3", "std_r3", "time_offset_imu_to_cam_s", "init_line_delay_s"):
        ds[k] = float(ds[k])
    ds["image_size"] = tuple(int(v) for v in ds["image_size"])
    return ds
