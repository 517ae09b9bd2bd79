"""Probe (not a test): wall-clock phases of the whole job through the C-ABI.  python tests/e2e_probe.py [config]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from openimucameracalibrator_b200 import _capi as capi, calibrator, synthetic as syn
cfg = syn.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 4]
ds = syn.make_dataset(cfg)
if os.environ.get("E2E_PINNED"):
    keep = []
    for k in ("uv", "point_ids", "accel", "gyro", "imu_t"):
        t = torch.from_numpy(np.ascontiguousarray(ds[k])).pin_memory(); keep.append(t); ds[k] = t.numpy()
lib = calibrator.load_library()
FLAGS = capi.FLAG_SPLINE | capi.FLAG_T_I_C
W, H = ds["image_size"]
rows = []
for rep in range(6):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    def lap(): torch.cuda.synchronize(); t.append(time.perf_counter())
    a = capi.CApi(lib, "icc_", 0); lap()
    a.set_camera(ds["model"], ds["intrinsics"], W, H); a.set_board_points(ds["board_xyzw"]); lap()
    a.set_frames(ds["frame_t"], ds["corner_offsets"], ds["point_ids"], ds["uv"], ds["q_wc"], ds["p_wc"]); lap()
    a.set_imu(ds["imu_t"], ds["accel"], ds["gyro"]); lap()
    a.batch_init_spline(ds["T_i_c_init"], ds["dt_so3_s"], ds["dt_r3_s"], ds["std_so3"], ds["std_r3"], ds["time_offset_imu_to_cam_s"], ds["init_line_delay_s"],
                        acc_bias=ds["acc_bias"], gyr_bias=ds["gyr_bias"], dispatch_fov=False); lap()
    a.set_known_gravity_dir(ds["gravity"]); lap()
    s = a.optimize(50, FLAGS); lap()
    T = a.get_T_i_c(); ld = a.get_line_delay(); lap()
    a.close(); lap()
    rows.append(np.diff(t) * 1e3)
    if rep == 0: print("iterations", s.iterations, "jac", s.seconds_jacobian, "solve", s.seconds_linear_solve, "total", s.seconds_total)
names = ["create", "camera+board", "set_frames", "set_imu", "batch_init", "gravity", "optimize", "getters", "close"]
r = np.array(rows)
for i, n in enumerate(names):
    print(f"{n:14s} " + " ".join(f"{x:7.3f}" for x in r[:, i]))
print("sum w/o close  " + " ".join(f"{x:7.3f}" for x in r[:, :-1].sum(1)))
