"""Developer probe: where the whole-job (e2e) wall time goes."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from openimucameracalibrator_b200 import _capi as capi, calibrator, synthetic as syn
F = capi.FLAG_SPLINE | capi.FLAG_T_I_C
ds = syn.make_dataset(syn.CONFIGS[4])
lib = calibrator.load_library()
for rep in range(6):
    t0 = time.perf_counter(); a = capi.CApi(lib, "icc_", 0); t1 = time.perf_counter()
    W, H = ds["image_size"]
    a.set_camera(ds["model"], ds["intrinsics"], W, H); a.set_board_points(ds["board_xyzw"])
    a.set_frames(ds["frame_t"], ds["corner_offsets"], ds["point_ids"], ds["uv"], ds["q_wc"], ds["p_wc"]); t1b = time.perf_counter()
    a.set_imu(ds["imu_t"], ds["accel"], ds["gyro"]); t2 = time.perf_counter()
    a.batch_init_spline(ds["T_i_c_init"], ds["dt_so3_s"], ds["dt_r3_s"], ds["std_so3"], ds["std_r3"], ds["time_offset_imu_to_cam_s"], ds["init_line_delay_s"], acc_bias=ds["acc_bias"], gyr_bias=ds["gyr_bias"]); t3 = time.perf_counter()
    a.set_known_gravity_dir(ds["gravity"]); t4 = time.perf_counter()
    s = a.optimize(50, F); t5 = time.perf_counter()
    T = a.get_T_i_c(); t6 = time.perf_counter(); a.close(); t7 = time.perf_counter()
    print(f"rep {rep}: create {1e3*(t1-t0):.2f}  set_frames {1e3*(t1b-t1):.2f}  set_imu {1e3*(t2-t1b):.2f}  batch_init {1e3*(t3-t2):.2f}  gravity {1e3*(t4-t3):.2f}  optimize {1e3*(t5-t4):.2f} (iters {s.iterations}, lm {1e3*s.seconds_total:.2f})  get {1e3*(t6-t5):.2f}  close {1e3*(t7-t6):.2f}  total {1e3*(t7-t0):.2f} ms")
