"""Python host mirror of OpenICC::core::ImuCameraCalibrator (include/OpenCameraCalibrator/core/imu_camera_calibrator.h:29-118)
on top of the CUDA C-ABI.  Same method names, argument meaning and error behaviour; there is no CPU fallback — a missing
`libicc_b200.so` or a missing GPU raises."""
from __future__ import annotations

import ctypes
import enum
import os

import numpy as np

from . import _capi
from ._capi import CApi, IccError

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def library_path() -> str:
    return os.path.join(_HERE, "libicc_b200.so")


def load_library() -> ctypes.CDLL:
    """Load the in-tree CUDA library; fails loudly when it has not been built (`python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise IccError(f"{path} is missing: build the CUDA extension first (__graft_entry__.build()); there is no CPU fallback")
        _LIB = ctypes.CDLL(path)
    return _LIB


class SplineOptimFlags(enum.IntFlag):
    """core/spline_trajectory_estimator.h:17-27"""
    POINTS = 1
    T_I_C = 2
    IMU_BIASES = 4
    IMU_INTRINSICS = 8
    GRAVITY_DIR = 16
    CAM_LINE_DELAY = 32
    SPLINE = 64
    ACC_BIAS = 128
    GYR_BIAS = 256


class ImuCameraCalibrator:
    """`BatchInitSpline` -> `Optimize` -> getters, as driven by continuous_time_imu_to_camera_calibration.cc:191-263."""

    def __init__(self, device: int = 0):
        self.api = CApi(load_library(), "icc_", int(device))
        self._cam_timestamps = None

    # -- problem data (what main() assembles into theia::Reconstruction / CameraTelemetryData) --------------------
    def SetCamera(self, model: int, intrinsics, image_width: int, image_height: int):
        self.api.set_camera(model, intrinsics, image_width, image_height)

    def SetBoardPoints(self, xyzw):
        self.api.set_board_points(xyzw)

    def SetViews(self, timestamps_s, corner_offsets, point_ids, uv, q_wc_xyzw, p_wc):
        self._cam_timestamps = np.sort(np.asarray(timestamps_s, dtype=np.float64))
        self.api.set_frames(timestamps_s, corner_offsets, point_ids, uv, q_wc_xyzw, p_wc)

    def SetTelemetry(self, timestamps_s, accelerometer, gyroscope):
        self.api.set_imu(timestamps_s, accelerometer, gyroscope)

    def SetShard(self, rank: int, world: int):
        self.api.set_shard(rank, world)

    # -- reference API ------------------------------------------------------------------------------------------------
    def BatchInitSpline(self, T_i_c_init, dt_so3, dt_r3, std_so3, std_r3, time_offset_imu_to_cam, initial_line_delay,
                        accl_intrinsics=(0, 0, 0, 1, 1, 1), gyro_intrinsics=(0, 0, 0, 0, 0, 0, 1, 1, 1), accl_bias=(0, 0, 0),
                        gyro_bias=(0, 0, 0), dispatch_fov=False):
        self.api.batch_init_spline(T_i_c_init, dt_so3, dt_r3, std_so3, std_r3, time_offset_imu_to_cam, initial_line_delay,
                                   accl_intrinsics, gyro_intrinsics, accl_bias, gyro_bias, dispatch_fov)

    def SetKnownGravityDir(self, gravity):
        self.api.set_known_gravity_dir(gravity)

    def Optimize(self, iterations: int, optim_flags: int) -> float:
        """Returns the mean reprojection error like the reference (imu_camera_calibrator.cc:163-168)."""
        self.last_summary = self.api.optimize(iterations, int(optim_flags))
        return self.last_summary.mean_reproj_error

    def GetCamTimestamps(self):
        return self._cam_timestamps

    def GetGyroMeasurements(self):
        t, _, g = self.api.imu_used()
        return t, g

    def GetAcclMeasurements(self):
        t, a, _ = self.api.imu_used()
        return t, a

    def GetCalibratedRSLineDelay(self) -> float:
        return self.api.get_line_delay()

    def GetT_i_c(self):
        return self.api.get_T_i_c()

    def GetGravity(self):
        return self.api.get_gravity()

    def EvalTrajectory(self, t_ns):
        """GetAngularVelocity / GetAcceleration / GetGyroBias / GetAcclBias / GetPose for many timestamps at once."""
        return self.api.eval_trajectory(t_ns)

    @classmethod
    def from_dataset(cls, ds: dict, device: int = 0, known_gravity: bool = True, shard=None) -> "ImuCameraCalibrator":
        self = cls(device)
        _capi.load_dataset(self.api, ds, known_gravity=known_gravity, shard=shard)
        self._cam_timestamps = np.sort(np.asarray(ds["frame_t"], dtype=np.float64))
        return self


class CameraCalibrator:
    """Python host mirror of OpenICC::core::CameraCalibrator (include/OpenCameraCalibrator/core/camera_calibrator.h:25-86,
    SURVEY.md §8(f) row f4): `CalibrateCameraFromJson` (src/core/camera_calibrator.cc:221-389) with the corner-file dictionary the
    reference's tools exchange, or `AddView` / `AddObservation` / `RunCalibration` (:78-219) with caller-made initial poses.  The
    bundle adjustment runs on the GPU behind `icc_calibrate_camera`; there is no CPU fallback."""

    def __init__(self, camera_model: str, optimize_board_pts: bool = False, device: int = 0):
        from . import camera_models as cm
        if camera_model not in cm.MODEL_IDS:
            raise IccError(f"unknown camera model {camera_model!r}")
        self.camera_model, self.model = camera_model, cm.MODEL_IDS[camera_model]
        self.optimize_board_pts = bool(optimize_board_pts)        # camera_calibrator.cc:207-216; refined points: self.api.get_board_points()
        self.api = CApi(load_library(), "icc_", int(device))
        self.grid_size, self.verbose = 0.04, False
        self._views, self._obs = [], {}
        self._image_size = None
        self.result = None

    def SetGridSize(self, grid_size: float = 0.04):
        self.grid_size = float(grid_size)

    def SetVerbose(self):
        self.verbose = True

    def SetBoardPoints(self, xyzw):
        self.api.set_board_points(xyzw)

    def AddView(self, initial_rotation, initial_position, initial_focal_length, initial_distortion, image_width, image_height, timestamp_s, group_id=0):
        """initial_rotation = R_cw (theia::Camera::SetOrientationFromRotationMatrix), initial_position = camera centre in the world."""
        from .synthetic import matrix_to_quat_xyzw
        R = np.asarray(initial_rotation, dtype=np.float64).reshape(3, 3)
        self._views.append(dict(q_wc=matrix_to_quat_xyzw(R.T[None])[0], p_wc=np.asarray(initial_position, dtype=np.float64), f=float(initial_focal_length),
                                k=float(initial_distortion), t=float(timestamp_s)))
        self._image_size = (int(image_width), int(image_height))
        self._obs[len(self._views) - 1] = []
        return len(self._views) - 1

    def AddObservation(self, view_id, object_point_id, corner) -> bool:
        if view_id not in self._obs:
            return False
        self._obs[view_id].append((int(object_point_id), float(corner[0]), float(corner[1])))
        return True

    def RunCalibration(self) -> bool:
        """Three-stage bundle adjustment from the views added so far (all of them: the grid filter belongs to CalibrateCameraFromJson)."""
        if not self._views:
            return False
        off, ids, uv = [0], [], []
        for v in range(len(self._views)):
            for pid, x, y in self._obs[v]:
                ids.append(pid); uv.append((x, y))
            off.append(len(ids))
        last = self._views[-1]                                     # the shared intrinsics keep the values of the last AddView (:101-113)
        W, H = self._image_size
        self.result = self.api.calibrate_camera(self.model, W, H, off, ids, np.array(uv).reshape(-1, 2), q_wc_init=np.array([v["q_wc"] for v in self._views]),
                                                p_wc_init=np.array([v["p_wc"] for v in self._views]), focal_length_init=last["f"], distortion_init=last["k"], grid_size=0.0,
                                                optimize_board_points=int(self.optimize_board_pts))
        return bool(self.result["summary"]["success"])

    def CalibrateCameraFromJson(self, scene_json: dict, output_path: str = "") -> bool:
        """scene_json: the corner file as a dictionary (io_formats.ubjson_loads of extract_board_to_json's output).  Writes
        `<output_path>.json` in the layout of io::write_camera_calibration when output_path is given."""
        import json
        from . import camera_models as cm
        pts = scene_json["scene_pts"]
        n = max(int(i) for i in pts) + 1
        board = np.zeros((n, 4)); board[:, 3] = 1.0
        for i, p in pts.items():
            board[int(i), :3] = p
        self.api.set_board_points(board)
        W, H = int(scene_json["image_width"]), int(scene_json["image_height"])
        off, ids, uv, self._timestamps_s = [0], [], [], []
        for key in sorted(scene_json["views"]):                    # nlohmann::json objects iterate in key order
            ip = scene_json["views"][key]["image_points"]
            for pid in sorted(ip):
                ids.append(int(pid)); uv.append(ip[pid])
            off.append(len(ids)); self._timestamps_s.append(float(key) * 1e-6)
        self.result = self.api.calibrate_camera(self.model, W, H, off, ids, np.array(uv, dtype=np.float64).reshape(-1, 2), grid_size=self.grid_size,
                                                optimize_board_points=int(self.optimize_board_pts))
        s = self.result["summary"]
        if not s["success"]:
            return False
        if output_path:
            k = self.result["intrinsics"]
            noskew = self.model in (cm.FOV, cm.DIVISION_UNDISTORTION)
            intr = {"skew": 0.0, "principal_pt_x": k[2] if noskew else k[3], "principal_pt_y": k[3] if noskew else k[4], "aspect_ratio": k[1], "focal_length": k[0]}
            extra = {cm.DIVISION_UNDISTORTION: {"div_undist_distortion": 4}, cm.DOUBLE_SPHERE: {"xi": 5, "alpha": 6}, cm.EXTENDED_UNIFIED: {"alpha": 5, "beta": 6},
                     cm.FISHEYE: {f"radial_distortion_{i + 1}": 5 + i for i in range(4)}, cm.FOV: {"radial_distortion_1": 4},
                     cm.PINHOLE_RADIAL_TANGENTIAL: {"radial_distortion_1": 5, "radial_distortion_2": 6, "radial_distortion_3": 7, "tangential_distortion_1": 8, "tangential_distortion_2": 9}}
            for name, idx in extra.get(self.model, {}).items():
                intr[name] = k[idx]
            doc = {"stabelized": False, "fps": scene_json["camera_fps"], "nr_calib_images": s["n_views_used"], "final_reproj_error": s["final_reproj_error"],
                   "image_width": W, "image_height": H, "intrinsic_type": self.camera_model, "intrinsics": {a: float(b) for a, b in intr.items()}}
            with open(output_path + ".json", "w") as f:
                json.dump(doc, f, indent=2)
        return True

    def PrintResult(self):
        from . import camera_models as cm
        k = self.result["intrinsics"]
        noskew = self.model in (cm.FOV, cm.DIVISION_UNDISTORTION)
        print(f"Focal Length:{k[0]}px Principal Point: {k[2] if noskew else k[3]}/{k[3] if noskew else k[4]}px.")
        if self.model == cm.DIVISION_UNDISTORTION:
            print(f"DIVISION_UNDISTORTION model: Distortion: {k[4]}")
        elif self.model == cm.DOUBLE_SPHERE:
            print(f"DOUBLE_SPHERE model: XI: {k[5]} ALPHA: {k[6]}")
        elif self.model == cm.EXTENDED_UNIFIED:
            print(f"EXTENDED_UNIFIED model: {k[5]} BETA: {k[6]}")
        elif self.model == cm.FISHEYE:
            print("FISHEYE model: " + " ".join(f"Radial distortion {i + 1}: {k[5 + i]}" for i in range(4)))
