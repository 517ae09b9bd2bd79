// Scalar pieces of the rotation / time-offset initialiser (icc_rotinit.cu) that run per thread: kept __host__ __device__ in a header so that
// tests/test_host_device_math.py can compile exactly this code for the CPU (nearest-sample rule of the reference, Eigen's slerp, the
// dominant eigenvector behind the closed-form rotation).
#pragma once
#include "icc_device_math.cuh"

namespace icc {

// FindClosestTimestamp (utils.cc:194-212) on sorted times: first strict minimum of |t - ts[i]|
ICC_HD int nearest_sorted(const double* __restrict__ ts, int n, double t, double& dist) {
  int lo = 0, hi = n;                    // first index with ts >= t
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (ts[mid] < t) lo = mid + 1; else hi = mid; }
  int idx;
  if (lo == 0) idx = 0;
  else if (lo == n) idx = n - 1;
  else idx = (fabs(t - ts[lo - 1]) <= fabs(t - ts[lo])) ? lo - 1 : lo;
  dist = fabs(t - ts[idx]);
  return idx;
}

ICC_HD double4 slerp4(double4 a, double4 b, double t) {            // Eigen::Quaternion::slerp (utils.cc:234)
  const double thresh = 1.0 - 2.220446049250313e-16;
  const double d = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w, ad = fabs(d);
  double s0, s1;
  if (ad >= thresh) { s0 = 1.0 - t; s1 = t; }
  else { const double th = acos(ad), sth = sin(th); s0 = sin((1.0 - t) * th) / sth; s1 = sin(t * th) / sth; }
  if (d < 0) s1 = -s1;
  return make_double4(s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z, s0 * a.w + s1 * b.w);
}

// largest eigenvector of the symmetric 4x4 matrix N by cyclic Jacobi rotations
ICC_HD void eig4_max(double (&A)[4][4], double (&qv)[4]) {
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 4; ++p) for (int q = p + 1; q < 4; ++q) off += A[p][q] * A[p][q];
    if (off < 1e-300) break;
    for (int p = 0; p < 4; ++p) for (int q = p + 1; q < 4; ++q) {
      if (fabs(A[p][q]) < 1e-300) continue;
      const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
      for (int k = 0; k < 4; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = cs * akp - sn * akq; A[k][q] = sn * akp + cs * akq; }
      for (int k = 0; k < 4; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = cs * apk - sn * aqk; A[q][k] = sn * apk + cs * aqk; }
      for (int k = 0; k < 4; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = cs * vkp - sn * vkq; V[k][q] = sn * vkp + cs * vkq; }
    }
  }
  int best = 0;
  for (int k = 1; k < 4; ++k) if (A[k][k] > A[best][best]) best = k;
  for (int k = 0; k < 4; ++k) qv[k] = V[k][best];
}

}  // namespace icc
