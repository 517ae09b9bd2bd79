"""Deterministic synthetic board + IMU sequences for the BASELINE.json configurations.

Everything the hot CLI reads from files (continuous_time_imu_to_camera_calibration.cc:91-199) is produced here as
arrays: board points, per-frame corner observations (rolling-shutter, the reference's own row-time model — SURVEY quirk
q2: row time `y * line_delay` is added to the *normalised* knot time), per-frame PnP-like pose priors, IMU streams with
bias + noise, spline weighting, initial T_i_c.  The same arrays drive the CUDA solver, the CPU oracle and the file
writers in `io_formats.py`.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import camera_models as cm


@dataclass
class SyntheticConfig:
    name: str
    model: int
    intrinsics: tuple
    n_frames: int
    grid: tuple            # (cols, rows) -> C = cols*rows corners
    imu_rate_hz: float
    seed: int
    fps: float = 30.0
    image_size: tuple = (960, 540)
    dt_so3_s: float = 0.05
    dt_r3_s: float = 0.05
    square_m: float = 0.021
    line_delay_truth: float = field(default=None)   # default 1/fps/height (app :186)
    line_delay_init_scale: float = 1.0
    time_offset_imu_to_cam_s: float = -0.0813
    corner_noise_px: float = 0.2
    pose_noise_m: float = 2e-3
    pose_noise_rad: float = np.deg2rad(0.2)
    tic_rot_perturb_rad: float = np.deg2rad(1.0)
    imu_noise: bool = True


# Truth intrinsics from the reference's README tables (Readme.md:31-39), Theia parameter order.
_F, _CX, _CY = 437.1, 489.1, 270.9
CONFIGS = {
    1: SyntheticConfig("cfg1_division_undistortion_60x35", cm.DIVISION_UNDISTORTION, (437.1, 1.0, 489.1, 270.9, -1.44e-6), 60, (7, 5), 200.0, 1235),
    2: SyntheticConfig("cfg2_fisheye_300x96", cm.FISHEYE, (435.5, 1.0, 0.0, 479.1, 274.5, 0.05, 0.07, -0.11, 0.05), 300, (12, 8), 200.0, 1236),
    3: SyntheticConfig("cfg3_double_sphere_rs_1000x96", cm.DOUBLE_SPHERE, (342.4, 1.0, 0.0, 472.6, 273.9, -0.215, 0.513), 1000, (12, 8), 400.0, 1237,
                       line_delay_init_scale=1.1),
    4: SyntheticConfig("cfg4_extended_unified_3000x144", cm.EXTENDED_UNIFIED, (438.0, 1.0, 0.0, 489.5, 272.0, 0.5115, 1.062), 3000, (16, 9), 1000.0, 1238),
}
# config 5: eight independent config-2-shaped sequences, models cycling
_CFG5_MODELS = [
    (cm.PINHOLE, (437.0, 1.0, 0.0, 489.0, 271.0, -0.05, 0.01)),
    (cm.FISHEYE, (435.5, 1.0, 0.0, 479.1, 274.5, 0.05, 0.07, -0.11, 0.05)),
    (cm.DIVISION_UNDISTORTION, (437.1, 1.0, 489.1, 270.9, -1.44e-6)),
    (cm.DOUBLE_SPHERE, (342.4, 1.0, 0.0, 472.6, 273.9, -0.215, 0.513)),
    (cm.EXTENDED_UNIFIED, (438.0, 1.0, 0.0, 489.5, 272.0, 0.5115, 1.062)),
    (cm.FOV, (437.0, 1.0, 489.0, 271.0, 0.9)),
    (cm.FISHEYE, (435.5, 1.0, 0.0, 479.1, 274.5, 0.05, 0.07, -0.11, 0.05)),
    (cm.EXTENDED_UNIFIED, (438.0, 1.0, 0.0, 489.5, 272.0, 0.5115, 1.062)),
]


def config5(k: int) -> SyntheticConfig:
    model, intr = _CFG5_MODELS[k % 8]
    return SyntheticConfig(f"cfg5_seq{k}_{cm.MODEL_NAMES[model].lower()}_300x96", model, intr, 300, (12, 8), 200.0, 1234 + 10 * k)


def tiny_config(model: int = cm.DIVISION_UNDISTORTION, intr=(437.1, 1.0, 489.1, 270.9, -1.44e-6), n_frames: int = 24,
                grid=(5, 4), imu_rate_hz: float = 100.0, seed: int = 7, **kw) -> SyntheticConfig:
    return SyntheticConfig(f"tiny_{cm.MODEL_NAMES[model].lower()}", model, tuple(intr), n_frames, grid, imu_rate_hz, seed, **kw)


# ---- small SO(3) helpers (numpy, batch) --------------------------------------------------------------------------

def _hat(v):
    z = np.zeros_like(v[..., 0])
    return np.stack([np.stack([z, -v[..., 2], v[..., 1]], -1), np.stack([v[..., 2], z, -v[..., 0]], -1),
                     np.stack([-v[..., 1], v[..., 0], z], -1)], -2)


def so3_exp(phi):
    phi = np.asarray(phi, dtype=np.float64)
    th = np.linalg.norm(phi, axis=-1)[..., None, None]
    K = _hat(phi)
    small = th < 1e-8
    ths = np.where(small, 1.0, th)
    A = np.where(small, 1.0 - th * th / 6.0, np.sin(ths) / ths)
    B = np.where(small, 0.5 - th * th / 24.0, (1.0 - np.cos(ths)) / (ths * ths))
    return np.eye(3) + A * K + B * (K @ K)


def so3_right_jacobian(phi):
    th = np.linalg.norm(phi, axis=-1)[..., None, None]
    K = _hat(phi)
    small = th < 1e-8
    ths = np.where(small, 1.0, th)
    A = np.where(small, 0.5 - th * th / 24.0, (1.0 - np.cos(ths)) / (ths * ths))
    B = np.where(small, 1.0 / 6.0 - th * th / 120.0, (ths - np.sin(ths)) / (ths ** 3))
    return np.eye(3) - A * K + B * (K @ K)


def quat_xyzw_to_matrix(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def matrix_to_quat_xyzw(R):
    """Batch rotation matrices (..., 3, 3) -> unit quaternions (..., 4) in x,y,z,w order with w >= 0."""
    R = np.asarray(R, dtype=np.float64)
    out = np.empty(R.shape[:-2] + (4,))
    flat_R = R.reshape(-1, 3, 3)
    flat_o = out.reshape(-1, 4)
    for i, M in enumerate(flat_R):
        t = np.trace(M)
        if t > 0:
            s = np.sqrt(t + 1.0) * 2
            q = np.array([(M[2, 1] - M[1, 2]) / s, (M[0, 2] - M[2, 0]) / s, (M[1, 0] - M[0, 1]) / s, 0.25 * s])
        elif M[0, 0] > M[1, 1] and M[0, 0] > M[2, 2]:
            s = np.sqrt(1.0 + M[0, 0] - M[1, 1] - M[2, 2]) * 2
            q = np.array([0.25 * s, (M[0, 1] + M[1, 0]) / s, (M[0, 2] + M[2, 0]) / s, (M[2, 1] - M[1, 2]) / s])
        elif M[1, 1] > M[2, 2]:
            s = np.sqrt(1.0 + M[1, 1] - M[0, 0] - M[2, 2]) * 2
            q = np.array([(M[0, 1] + M[1, 0]) / s, 0.25 * s, (M[1, 2] + M[2, 1]) / s, (M[0, 2] - M[2, 0]) / s])
        else:
            s = np.sqrt(1.0 + M[2, 2] - M[0, 0] - M[1, 1]) * 2
            q = np.array([(M[0, 2] + M[2, 0]) / s, (M[1, 2] + M[2, 1]) / s, 0.25 * s, (M[1, 0] - M[0, 1]) / s])
        if q[3] < 0:
            q = -q
        flat_o[i] = q / np.linalg.norm(q)
    return out


class Trajectory:
    """Analytic IMU trajectory: R_wi(t) = R_base exp(phi(t)), p_wi(t) = p_base + s(t); sums of three sinusoids per axis."""

    def __init__(self, rng, R_base, p_base, rot_amp=0.15, pos_amp=0.10):
        self.R_base, self.p_base = R_base, p_base
        self.f_rot = rng.uniform(0.2, 1.5, size=(3, 3))
        self.f_pos = rng.uniform(0.2, 1.5, size=(3, 3))
        self.ph_rot = rng.uniform(0, 2 * np.pi, size=(3, 3))
        self.ph_pos = rng.uniform(0, 2 * np.pi, size=(3, 3))
        self.a_rot = rng.uniform(0.5, 1.0, size=(3, 3)) * rot_amp / 3.0
        self.a_pos = rng.uniform(0.5, 1.0, size=(3, 3)) * pos_amp / 3.0

    @staticmethod
    def _sin(t, a, f, ph, order):
        w = 2 * np.pi * f
        arg = w[None] * np.asarray(t)[..., None, None] + ph[None]
        if order == 0:
            v = a * np.sin(arg)
        elif order == 1:
            v = a * w * np.cos(arg)
        else:
            v = -a * w * w * np.sin(arg)
        return v.sum(-1)

    def phi(self, t, order=0):
        return self._sin(t, self.a_rot, self.f_rot, self.ph_rot, order)

    def R_wi(self, t):
        return self.R_base @ so3_exp(self.phi(t))

    def p_wi(self, t, order=0):
        s = self._sin(t, self.a_pos, self.f_pos, self.ph_pos, order)
        return s + self.p_base if order == 0 else s

    def omega_body(self, t):
        phi, dphi = self.phi(t), self.phi(t, 1)
        return (so3_right_jacobian(phi) @ dphi[..., None])[..., 0]


def make_dataset(cfg: SyntheticConfig) -> dict:
    rng = np.random.default_rng(cfg.seed)
    W, H = cfg.image_size
    ld_true = cfg.line_delay_truth if cfg.line_delay_truth is not None else 1.0 / cfg.fps / H
    # board: planar grid in z = 0, centred, ids row-major; homogeneous points
    cols, rows = cfg.grid
    gx, gy = np.meshgrid((np.arange(cols) - (cols - 1) / 2) * cfg.square_m, (np.arange(rows) - (rows - 1) / 2) * cfg.square_m)
    board = np.stack([gx.ravel(), gy.ravel(), np.zeros(cols * rows), np.ones(cols * rows)], -1)
    C = board.shape[0]
    # truth extrinsics: T_i_c (camera pose in the IMU frame), README dataset 1
    q_ic = np.array([-0.007, -0.708, 0.706, 0.005]); q_ic /= np.linalg.norm(q_ic)
    R_ic = quat_xyzw_to_matrix(q_ic)
    t_ic = np.array([0.007, -0.022, 0.001])
    # camera looks down on the board from ~0.5 m; IMU base pose follows from T_wc = T_wi * T_i_c
    R_wc_base = np.diag([1.0, -1.0, -1.0])
    p_wc_base = np.array([0.0, 0.0, 0.5])
    R_wi_base = R_wc_base @ R_ic.T
    p_wi_base = p_wc_base - R_wi_base @ t_ic
    traj = Trajectory(rng, R_wi_base, p_wi_base)
    gravity = np.array([0.0, 0.0, 9.81])

    frame_t = np.arange(cfg.n_frames) / cfg.fps
    # rolling shutter observation times in the reference's model: physical offset = y * ld * dt (dt_so3 == dt_r3 here)
    dt_phys = cfg.dt_so3_s
    uv = np.zeros((cfg.n_frames, C, 2))
    y_row = np.full((cfg.n_frames, C), H / 2.0)
    X = board[:, :3]
    for _ in range(4):
        t_obs = frame_t[:, None] + y_row * ld_true * dt_phys
        R = traj.R_wi(t_obs.ravel()).reshape(cfg.n_frames, C, 3, 3)
        p = traj.p_wi(t_obs.ravel()).reshape(cfg.n_frames, C, 3)
        q_imu = np.einsum("fcji,fcj->fci", R, X[None] - p)           # R_wi^T (X - p_wi)
        p_cam = np.einsum("ji,fcj->fci", R_ic, q_imu - t_ic)            # R_ic^T (q - t_ic)
        uv, valid = cm.project(cfg.model, cfg.intrinsics, p_cam)
        y_row = uv[..., 1]
    assert valid.all(), "synthetic trajectory left the camera model's valid domain"
    uv_noisy = uv + rng.normal(0.0, cfg.corner_noise_px, size=uv.shape)

    # per-frame pose priors (PnP-like): T_wc at the frame timestamp + noise
    R_wi_f = traj.R_wi(frame_t)
    p_wi_f = traj.p_wi(frame_t)
    R_wc = R_wi_f @ R_ic
    p_wc = p_wi_f + (R_wi_f @ t_ic)
    R_wc_n = R_wc @ so3_exp(rng.normal(0.0, cfg.pose_noise_rad, size=(cfg.n_frames, 3)))
    p_wc_n = p_wc + rng.normal(0.0, cfg.pose_noise_m, size=p_wc.shape)
    q_wc = matrix_to_quat_xyzw(R_wc_n)

    # IMU streams on the IMU clock; camera time = imu time + offset
    duration = frame_t[-1]
    n_imu = int(np.floor((duration + 0.5) * cfg.imu_rate_hz)) + 1
    t_cam_clock = np.arange(n_imu) / cfg.imu_rate_hz - 0.25
    R_i = traj.R_wi(t_cam_clock)
    acc = np.einsum("nji,nj->ni", R_i, traj.p_wi(t_cam_clock, 2) + gravity)
    gyr = traj.omega_body(t_cam_clock)
    b_a = np.array([0.05, -0.03, 0.02]); b_g = np.array([2e-3, -1e-3, 5e-4])
    sig_a = 1.5e-2 * np.sqrt(cfg.imu_rate_hz); sig_g = 1.1e-3 * np.sqrt(cfg.imu_rate_hz)
    if cfg.imu_noise:
        acc = acc + rng.normal(0.0, sig_a, size=acc.shape)
        gyr = gyr + rng.normal(0.0, sig_g, size=gyr.shape)
    acc = acc + b_a
    gyr = gyr + b_g
    imu_t = t_cam_clock - cfg.time_offset_imu_to_cam_s

    # initial T_i_c: rotation perturbed, translation zero (app :170, quirk q8)
    axis = rng.normal(size=3); axis /= np.linalg.norm(axis)
    R_ic_init = R_ic @ so3_exp(axis * cfg.tic_rot_perturb_rad)
    q_ic_init = matrix_to_quat_xyzw(R_ic_init)
    T_ic_init = np.concatenate([q_ic_init, np.zeros(3)])

    offsets = (np.arange(cfg.n_frames + 1) * C).astype(np.int32)
    point_ids = np.tile(np.arange(C, dtype=np.int32), cfg.n_frames)
    return dict(
        name=cfg.name, config=cfg, model=cfg.model, intrinsics=np.asarray(cfg.intrinsics, dtype=np.float64), image_size=(W, H), fps=cfg.fps,
        board_xyzw=board, frame_t=frame_t, corner_offsets=offsets, point_ids=point_ids, uv=uv_noisy.reshape(-1, 2).copy(),
        q_wc=q_wc, p_wc=p_wc_n, imu_t=imu_t, accel=acc, gyro=gyr,
        dt_so3_s=cfg.dt_so3_s, dt_r3_s=cfg.dt_r3_s, std_so3=sig_g if cfg.imu_noise else 1.0, std_r3=sig_a if cfg.imu_noise else 1.0,
        time_offset_imu_to_cam_s=cfg.time_offset_imu_to_cam_s, init_line_delay_s=ld_true * cfg.line_delay_init_scale,
        T_i_c_init=T_ic_init, acc_bias=b_a, gyr_bias=b_g, gravity=gravity,
        truth=dict(T_i_c=np.concatenate([q_ic, t_ic]), line_delay=ld_true, gravity=gravity),
    )
