import numpy as np

from openimucameracalibrator_b200 import _capi as capi

F_STAGE1 = capi.FLAG_SPLINE | capi.FLAG_T_I_C
F_STAGE2 = capi.FLAG_CAM_LINE_DELAY
F_ALL = F_STAGE1 | capi.FLAG_GRAVITY_DIR | capi.FLAG_CAM_LINE_DELAY | capi.FLAG_IMU_BIASES


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if a.size else 0.0


def qmul(a, b):
    ax, ay, az, aw = a; bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def qexp(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.array([*(0.5 * w), 1.0])
    return np.array([*(np.sin(th / 2) / th * w), np.cos(th / 2)])


def qmat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def plus_homog4(x, d):
    """ceres::HomogeneousVectorParameterization(4)::Plus (Hartley & Zisserman A6.9.2-3), NumPy statement."""
    x = np.asarray(x, dtype=np.float64); d = np.asarray(d, dtype=np.float64)
    nd = np.linalg.norm(d)
    if nd == 0.0:
        return x.copy()
    y = np.concatenate([0.5 * np.sin(0.5 * nd) / (0.5 * nd) * d, [np.cos(0.5 * nd)]])
    sigma = float(x[:3] @ x[:3]); v = np.array([*x[:3], 1.0]); beta = 0.0
    if sigma <= np.finfo(float).eps:
        beta = 2.0 if x[3] < 0 else 0.0
    else:
        mu = np.sqrt(x[3] ** 2 + sigma)
        vp = x[3] - mu if x[3] <= 0 else -sigma / (x[3] + mu)
        beta = 2.0 * vp * vp / (sigma + vp * vp); v[:3] /= vp
    return np.linalg.norm(x) * (y - v * (beta * (v @ y)))


class TangentWalker:
    """Applies a canonical-order tangent vector to a solver through its public setters (first order for T_i_c): used for
    finite-difference checks of gradients on either implementation."""

    def __init__(self, api, flags):
        self.api, self.flags = api, flags
        self.so3, self.r3, self.ba, self.bg = api.get_knots()
        self.T, self.ld, self.g = api.get_T_i_c(), api.get_line_delay(), api.get_gravity()
        self.pts = api.get_board_points() if flags & capi.FLAG_POINTS else None
        self.n = api.num_tangent(flags)

    def apply(self, d):
        f, off = self.flags, 0
        so3, r3, ba, bg, T, ld, g = self.so3.copy(), self.r3.copy(), self.ba.copy(), self.bg.copy(), self.T.copy(), self.ld, self.g.copy()
        if f & capi.FLAG_SPLINE:
            for i in range(len(so3)):
                so3[i] = qmul(self.so3[i], qexp(d[off + 3 * i: off + 3 * i + 3]))
            off += 3 * len(so3)
            r3 = self.r3 + d[off: off + r3.size].reshape(-1, 3); off += r3.size
        if f & capi.FLAG_T_I_C:
            ups, om = d[off: off + 3], d[off + 3: off + 6]; off += 6
            T = np.concatenate([qmul(self.T[:4], qexp(om)), self.T[4:] + qmat(self.T[:4]) @ ups])
        if f & capi.FLAG_GRAVITY_DIR:
            g = self.g + d[off: off + 3]; off += 3
        if (f & capi.FLAG_CAM_LINE_DELAY) and self.ld != 0.0:
            ld = self.ld + d[off]; off += 1
        if f & (capi.FLAG_IMU_BIASES | capi.FLAG_ACC_BIAS):
            ba = self.ba + d[off: off + ba.size].reshape(-1, 3); off += ba.size
        if f & (capi.FLAG_IMU_BIASES | capi.FLAG_GYR_BIAS):
            bg = self.bg + d[off: off + bg.size].reshape(-1, 3); off += bg.size
        if f & capi.FLAG_POINTS:           # board points: last canonical block, HomogeneousVectorParameterization(4)
            pts = np.array([plus_homog4(x, d[off + 3 * i: off + 3 * i + 3]) for i, x in enumerate(self.pts)]); off += 3 * len(self.pts)
            self.api.set_board_points(pts)
        assert off == self.n
        self.api.set_knots(so3, r3, ba, bg); self.api.set_T_i_c(T); self.api.set_line_delay(ld); self.api.set_known_gravity_dir(g)

    def cost(self, d):
        self.apply(d)
        return self.api.evaluate(self.flags, residuals=False, gradient=False)[0]

    def restore(self):
        self.apply(np.zeros(self.n))
