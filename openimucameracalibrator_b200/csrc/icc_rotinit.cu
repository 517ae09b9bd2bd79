// IMU-to-camera rotation + time-offset initialiser (sm_100a): the step that produces `gyro_to_cam_calibration.json`, i.e. the
// T_i_c_init rotation and the time offset the hot CLI starts from.
//
// Replaces what the reference runs serially on the CPU in
//   ImuToCameraRotationEstimator::EstimateCameraImuRotation   src/core/imu_to_camera_rotation_estimator.cc:116-274
//   ImuToCameraRotationEstimator::SolveClosedForm             src/core/imu_to_camera_rotation_estimator.cc:39-114
//   utils::InterpolateQuaternions / InterpolateVector3d / FindClosestTimestamp     src/utils/utils.cc:194-261
// The reference's FindClosestTimestamp is a linear scan per query (O(N^2) per objective evaluation, 40+ evaluations); on the
// sorted, distinct timestamps it is given, bisection + comparison of the two neighbours finds the same index.
//
// Mapping to the machine: every stage is data parallel over the IMU samples (one thread per sample): quaternion interpolation
// to the IMU rate, quaternion finite differences -> visual angular velocity, outlier hold, 15-tap moving averages; one
// evaluation of the alignment objective at a candidate time offset = a grid-wide reduction of 15 sums (means + 3x3
// cross-covariance), a single-thread closed-form rotation (Horn's quaternion method on the 3x3 covariance), and a second
// grid-wide reduction of the robust error.  The golden-section search keeps its state on the device: both candidates of an
// iteration are evaluated by the same launches and a one-thread kernel shrinks the bracket, so the host only enqueues launches
// and reads the result once at the end.
#include "icc_device_math.cuh"
#include "icc_kernels.h"

#include "icc_rotinit_math.cuh"

#include <cmath>

namespace icc {

void count_launch();

namespace {

constexpr double kHuberK = 1.345, kHuberK2 = kHuberK * kHuberK;   // imu_to_camera_rotation_estimator.cc:36-37

// InterpolateQuaternions (utils.cc:220-240): nearest old sample, slerp towards its successor by dist / (t[n+1] - t[n])
__global__ void interp_quat_kernel(int n_old, const double* __restrict__ t_old, const double4* __restrict__ q_old, int n_new, const double* __restrict__ t_new,
                                   double4* __restrict__ q_new) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_new) return;
  double dist;
  const int k = nearest_sorted(t_old, n_old, t_new[i], dist);
  q_new[i] = k < n_old - 1 ? slerp4(q_old[k], q_old[k + 1], dist / (t_old[k + 1] - t_old[k])) : q_old[k];
}

// visual angular velocity from quaternion finite differences (:178-207); flag = component beyond 2 pi
__global__ void angvel_kernel(int n, const double4* __restrict__ q, double dt_imu, double* __restrict__ w_raw, unsigned char* __restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = i < n - 1 ? i : n - 2;             // the last difference is duplicated (:188)
  double4 d = make_double4(0, 0, 0, 0);
  if (n >= 2) d = make_double4(q[j + 1].x - q[j].x, q[j + 1].y - q[j].y, q[j + 1].z - q[j].z, q[j + 1].w - q[j].w);
  const double4 qi = q[i];
  const double n2 = qi.x * qi.x + qi.y * qi.y + qi.z * qi.z + qi.w * qi.w;
  const Q4 inv = q4(-qi.x / n2, -qi.y / n2, -qi.z / n2, qi.w / n2);          // Eigen inverse = conjugate / squaredNorm
  const Q4 a = qmul(q4(d.x, d.y, d.z, d.w), inv);
  const double s = -2.0 / dt_imu;
  const double wx = s * a.x, wy = s * a.y, wz = s * a.z;
  w_raw[3 * i] = wx; w_raw[3 * i + 1] = wy; w_raw[3 * i + 2] = wz;
  const double lim = 2.0 * 3.14159265358979323846;
  bad[i] = (fabs(wx) > lim || fabs(wy) > lim || fabs(wz) > lim) ? 1 : 0;
}

// suppress extreme velocities (:196-205): a flagged sample repeats its (already repaired) predecessor, or is zero for i <= 1
__global__ void hold_kernel(int n, const double* __restrict__ w_raw, const unsigned char* __restrict__ bad, double* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int j = i;
  while (bad[j] && j > 1) --j;
  const bool zero = bad[j] != 0;                   // reached i <= 1 still flagged
  w[3 * i] = zero ? 0.0 : w_raw[3 * j]; w[3 * i + 1] = zero ? 0.0 : w_raw[3 * j + 1]; w[3 * i + 2] = zero ? 0.0 : w_raw[3 * j + 2];
}

// SimpleMovingAverage(15) (utils/moving_average.h): mean of the last min(i + 1, 15) samples
__global__ void smooth_kernel(int n, const double* __restrict__ x, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k0 = i >= 14 ? i - 14 : 0;
  double s0 = 0, s1 = 0, s2 = 0;
  for (int k = k0; k <= i; ++k) { s0 += x[3 * k]; s1 += x[3 * k + 1]; s2 += x[3 * k + 2]; }
  const double inv = 1.0 / (double)(i - k0 + 1);
  y[3 * i] = s0 * inv; y[3 * i + 1] = s1 * inv; y[3 * i + 2] = s2 * inv;
}

// ---- golden-section state on the device -------------------------------------------------------------------------------------
// sums[c][0..2] = sum vis, [3..5] = sum imu, [6..14] = sum imu_a vis_b (row a, column b), [15] = robust error
ICC_D double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) { t = lane < nw ? red[lane] : 0.0; for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o); }
  return t;                                        // valid in warp 0
}

// InterpolateVector3d(time_with_offset, timestamps, angVis) (:47-54, utils.cc:242-261) for sample i at offset td.
// The reference reads one element past the end when the nearest sample is the last one (undefined behaviour); here the last
// sample is returned as is.
ICC_D V3 shifted_vis(const RotInitProblem& Q, int i, double td) {
  const double ti = Q.t[i];
  double dist;
  int lo = 0, hi = Q.n;                            // nearest of (t[j] - td) to t[i]
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (Q.t[mid] - td < ti) lo = mid + 1; else hi = mid; }
  int k;
  if (lo == 0) k = 0;
  else if (lo == Q.n) k = Q.n - 1;
  else k = (fabs(ti - (Q.t[lo - 1] - td)) <= fabs(ti - (Q.t[lo] - td))) ? lo - 1 : lo;
  dist = fabs(ti - (Q.t[k] - td));
  const V3 v0 = v3(Q.vis[3 * k], Q.vis[3 * k + 1], Q.vis[3 * k + 2]);
  if (k + 1 >= Q.n) return v0;
  const double f = dist / ((Q.t[k + 1] - td) - (Q.t[k] - td));
  const V3 v1 = v3(Q.vis[3 * k + 3], Q.vis[3 * k + 4], Q.vis[3 * k + 5]);
  return v3((1.0 - f) * v0.x + f * v1.x, (1.0 - f) * v0.y + f * v1.y, (1.0 - f) * v0.z + f * v1.z);   // lerp3d (utils.cc:214-218)
}

__global__ void __launch_bounds__(256) sums_kernel(RotInitProblem Q) {
  __shared__ double red[8];
  if (Q.state->done) return;
  const int c = blockIdx.y;
  const double td = c == 0 ? Q.state->c : Q.state->d;
  double s[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) s[k] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Q.n; i += gridDim.x * blockDim.x) {
    const V3 v = shifted_vis(Q, i, td);
    Q.vis_shift[(size_t)c * 3 * Q.n + 3 * i] = v.x; Q.vis_shift[(size_t)c * 3 * Q.n + 3 * i + 1] = v.y; Q.vis_shift[(size_t)c * 3 * Q.n + 3 * i + 2] = v.z;
    const double a0 = Q.imu[3 * i], a1 = Q.imu[3 * i + 1], a2 = Q.imu[3 * i + 2];
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += a0; s[4] += a1; s[5] += a2;
    s[6] += a0 * v.x; s[7] += a0 * v.y; s[8] += a0 * v.z; s[9] += a1 * v.x; s[10] += a1 * v.y; s[11] += a1 * v.z; s[12] += a2 * v.x; s[13] += a2 * v.y; s[14] += a2 * v.z;
  }
#pragma unroll
  for (int k = 0; k < 15; ++k) { const double t = block_sum(s[k], red); if (threadIdx.x == 0) atomicAdd(&Q.state->sums[c][k], t); }
}

// closed-form rotation (:56-87): Rs = V C U^T of the SVD of P^T Q  ==  the proper rotation maximising tr(R P^T Q), obtained
// here as the dominant eigenvector of Horn's 4x4 matrix (no SVD needed); bias = mean_vis - Rs mean_imu when enabled
__global__ void solve_kernel(RotInitProblem Q) {
  RotInitState* S = Q.state;
  if (S->done) return;
  const int c = threadIdx.x;
  if (c >= 2) return;
  const double n = (double)Q.n;
  const double* s = S->sums[c];
  const double mv[3] = {s[0] / n, s[1] / n, s[2] / n}, mi[3] = {s[3] / n, s[4] / n, s[5] / n};
  double M[3][3];                                   // M[a][b] = sum (imu_a - mean)(vis_b - mean)
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) M[a][b] = s[6 + 3 * a + b] - n * mi[a] * mv[b];
  double N4[4][4] = {
      {M[0][0] + M[1][1] + M[2][2], M[1][2] - M[2][1], M[2][0] - M[0][2], M[0][1] - M[1][0]},
      {M[1][2] - M[2][1], M[0][0] - M[1][1] - M[2][2], M[0][1] + M[1][0], M[2][0] + M[0][2]},
      {M[2][0] - M[0][2], M[0][1] + M[1][0], -M[0][0] + M[1][1] - M[2][2], M[1][2] + M[2][1]},
      {M[0][1] - M[1][0], M[2][0] + M[0][2], M[1][2] + M[2][1], -M[0][0] - M[1][1] + M[2][2]}};
  double qv[4];
  eig4_max(N4, qv);                                 // (w, x, y, z) of the rotation taking imu to vis
  const Q4 q = qnormalized(q4(qv[1], qv[2], qv[3], qv[0]));
  const M3 R = qmat(q);
  for (int k = 0; k < 9; ++k) S->R[c][k] = R.m[k];
  const V3 rm = mul(R, v3(mi[0], mi[1], mi[2]));
  const bool est = Q.estimate_bias != 0;
  S->bias[c][0] = est ? mv[0] - rm.x : 0.0; S->bias[c][1] = est ? mv[1] - rm.y : 0.0; S->bias[c][2] = est ? mv[2] - rm.z : 0.0;
}

__global__ void __launch_bounds__(256) error_kernel(RotInitProblem Q) {
  __shared__ double red[8];
  const RotInitState* S = Q.state;
  if (S->done) return;
  const int c = blockIdx.y;
  M3 R; for (int k = 0; k < 9; ++k) R.m[k] = S->R[c][k];
  const V3 b = v3(S->bias[c][0], S->bias[c][1], S->bias[c][2]);
  double e = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Q.n; i += gridDim.x * blockDim.x) {
    const double* vs = Q.vis_shift + (size_t)c * 3 * Q.n + 3 * i;
    const V3 D = v3(vs[0], vs[1], vs[2]) - (mul(R, v3(Q.imu[3 * i], Q.imu[3 * i + 1], Q.imu[3 * i + 2])) + b);
    const double err = dot(D, D);
    e += err > kHuberK ? 2.0 * kHuberK * sqrt(err) - kHuberK2 : err;       // (:100-107: the SQUARED norm is compared with k)
  }
  const double t = block_sum(e, red);
  if (threadIdx.x == 0) atomicAdd(&Q.state->sums[c][15], t);
}

// one golden-section step (:232-257)
__global__ void golden_kernel(RotInitProblem Q) {
  RotInitState* S = Q.state;
  if (S->done) return;
  const double fc = S->sums[0][15], fd = S->sums[1][15];
  const int keep = fc < fd ? 0 : 1;
  if (keep == 0) S->b = S->d; else S->a = S->c;
  for (int k = 0; k < 9; ++k) S->R_best[k] = S->R[keep][k];
  if (Q.estimate_bias) for (int k = 0; k < 3; ++k) S->bias_best[k] = S->bias[keep][k];
  S->error = keep == 0 ? fc : fd;
  const double g = (1.0 + sqrt(5.0)) / 2.0;
  S->c = S->b - (S->b - S->a) / g; S->d = S->a + (S->b - S->a) / g;
  S->iterations += 1;
  for (int c = 0; c < 2; ++c) for (int k = 0; k < 16; ++k) S->sums[c][k] = 0.0;
  if (!(fabs(S->c - S->d) > Q.tolerance)) S->done = 1;
}

__global__ void init_state_kernel(RotInitProblem Q, double max_offset) {
  RotInitState* S = Q.state;
  const double g = (1.0 + sqrt(5.0)) / 2.0;
  S->a = -max_offset; S->b = max_offset;
  S->c = S->b - (S->b - S->a) / g; S->d = S->a + (S->b - S->a) / g;
  S->iterations = 0; S->error = 0.0;
  for (int c = 0; c < 2; ++c) { for (int k = 0; k < 16; ++k) S->sums[c][k] = 0.0; for (int k = 0; k < 9; ++k) S->R[c][k] = (k % 4 == 0) ? 1.0 : 0.0; for (int k = 0; k < 3; ++k) S->bias[c][k] = 0.0; }
  for (int k = 0; k < 9; ++k) S->R_best[k] = (k % 4 == 0) ? 1.0 : 0.0;
  for (int k = 0; k < 3; ++k) S->bias_best[k] = Q.bias_in[k];
  S->done = !(fabs(S->c - S->d) > Q.tolerance) ? 1 : 0;
}

}  // namespace

void launch_interp_quat(int n_old, const double* t_old, const double4* q_old, int n_new, const double* t_new, double4* q_new, cudaStream_t st) {
  if (n_new <= 0) return;
  interp_quat_kernel<<<(n_new + 255) / 256, 256, 0, st>>>(n_old, t_old, q_old, n_new, t_new, q_new); count_launch();
}

void launch_visual_angular_velocity(int n, const double4* q, double dt_imu, double* w_raw, unsigned char* bad, double* w_held, double* w_smooth, const double* imu, double* imu_smooth, cudaStream_t st) {
  if (n <= 0) return;
  const int g = (n + 255) / 256;
  angvel_kernel<<<g, 256, 0, st>>>(n, q, dt_imu, w_raw, bad); count_launch();
  hold_kernel<<<g, 256, 0, st>>>(n, w_raw, bad, w_held); count_launch();
  smooth_kernel<<<g, 256, 0, st>>>(n, w_held, w_smooth); count_launch();
  smooth_kernel<<<g, 256, 0, st>>>(n, imu, imu_smooth); count_launch();
}

void launch_golden_section(const RotInitProblem& Q, double max_offset, int max_iterations, int sm_count, cudaStream_t st) {
  init_state_kernel<<<1, 1, 0, st>>>(Q, max_offset); count_launch();
  int gx = (Q.n + 255) / 256; if (gx > 4 * sm_count) gx = 4 * sm_count; if (gx < 1) gx = 1;
  for (int it = 0; it < max_iterations; ++it) {
    sums_kernel<<<dim3(gx, 2), 256, 0, st>>>(Q); count_launch();
    solve_kernel<<<1, 32, 0, st>>>(Q); count_launch();
    error_kernel<<<dim3(gx, 2), 256, 0, st>>>(Q); count_launch();
    golden_kernel<<<1, 1, 0, st>>>(Q); count_launch();
  }
}

}  // namespace icc
