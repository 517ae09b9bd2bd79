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
