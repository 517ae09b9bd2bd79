// ceres::HomogeneousVectorParameterization(4) for one board point (Ceres 2.1 local_parameterization.cc, restated from the published
// algorithm, Hartley & Zisserman A6.9.2-3): Householder vector of x, the 4 x 3 local Jacobian |x| / 2 x (first three columns of H), Plus.
// __host__ __device__ so that tests/test_host_device_math.py compiles exactly this code for the CPU (against the NumPy statement of
// tests/helpers.py and finite differences); icc_points.cu inlines it unchanged.
#pragma once
#include "icc_device_math.cuh"

namespace icc {

// v (v[3] = 1) and beta with (I - beta v v^T) x = |x| e_4
ICC_HD void householder4(const double x[4], double (&v)[4], double& beta) {
  const double sigma = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
  v[0] = x[0]; v[1] = x[1]; v[2] = x[2]; v[3] = 1.0; beta = 0.0;
  if (sigma <= 2.220446049250313e-16) { if (x[3] < 0.0) beta = 2.0; return; }
  const double mu = sqrt(x[3] * x[3] + sigma);
  const double vp = x[3] <= 0.0 ? x[3] - mu : -sigma / (x[3] + mu);
  beta = 2.0 * vp * vp / (sigma + vp * vp);
  v[0] /= vp; v[1] /= vp; v[2] /= vp;
}

// de-homogenised point (w = 1) and the local Jacobian d x / d delta at delta = 0, row-major 4 x 3
ICC_HD void points_prepare(const double x[4], double board[4], double jac12[12]) {
  const double iw = 1.0 / x[3];
  board[0] = x[0] * iw; board[1] = x[1] * iw; board[2] = x[2] * iw; board[3] = 1.0;   // hnormalized(T^-1 X_h) == T^-1 (X / w)   (residuals.h:357-362)
  double v[4], beta; householder4(x, v, beta);
  const double n = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
  for (int k = 0; k < 4; ++k) for (int i = 0; i < 3; ++i) jac12[3 * k + i] = n * (-0.5 * beta * v[i] * v[k] + (k == i ? 0.5 : 0.0));
}

// x_plus = |x| H(x) [sin(|d| / 2) d / |d| ; cos(|d| / 2)]  (x itself for d = 0)
ICC_HD void points_plus(const double x[4], const double d[3], double out[4]) {
  const double nd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (nd == 0.0) { out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = x[3]; return; }
  const double hh = 0.5 * nd, sbd = sin(hh) / hh;
  const double y[4] = {0.5 * sbd * d[0], 0.5 * sbd * d[1], 0.5 * sbd * d[2], cos(hh)};
  double v[4], beta; householder4(x, v, beta);
  const double nx = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
  const double vy = v[0] * y[0] + v[1] * y[1] + v[2] * y[2] + v[3] * y[3];
  for (int k = 0; k < 4; ++k) out[k] = nx * (y[k] - v[k] * (beta * vy));
}

}  // namespace icc
