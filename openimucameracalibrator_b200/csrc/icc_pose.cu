// Per-view board pose estimation (sm_100a): the step that fills the pose dataset the hot CLI consumes.
//
// Replaces, for every view in parallel, what the reference runs serially on the CPU in
//   PoseEstimator::EstimatePosesFromJson      src/core/pose_estimator.cc:92-191
//   PoseEstimator::EstimatePosePinhole        src/core/pose_estimator.cc:54-90
// i.e. theia::Camera::PixelToNormalizedCoordinates per corner, a calibrated absolute pose (RANSAC PnP, squared normalised
// reprojection error threshold, >= 6 inliers) and theia::BundleAdjustView (pose-only, Huber 1.345) on the inliers.
//
// Mapping to the machine: kernel 1 un-projects one corner per thread by Gauss-Newton on the forward model (the closed-form
// 2x3 projection Jacobians of icc_camera.cuh; one code path for all seven camera models, residual < 1e-13 px); kernel 2 gives
// one WARP per view: lanes stride over the corners, the 8x8 (homography) and 6x6 (pose) normal equations are reduced with warp
// shuffles and solved redundantly in registers by every lane (no shared memory, no divergence).  The calibration board is
// planar, so the initial pose comes from the normalised DLT homography; the refinement minimises the same cost as the
// reference's BundleAdjustView (normalised pinhole reprojection error, Huber 1.345) with Levenberg-Marquardt to a tighter
// tolerance than Ceres' defaults, so both land on the same optimum.
#include "icc_camera.cuh"
#include "icc_kernels.h"
#include "icc_small_linalg.cuh"

#include <cmath>

namespace icc {

void count_launch();

namespace {

ICC_D double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct Intr { double k[10]; };

__global__ void unproject_kernel(int model, Intr K, int n, const double2* __restrict__ uv, double2* __restrict__ xy, int* __restrict__ ok) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = 0.0, y = 0.0;
  const bool good = unproject_gn(model, K.k, uv[i].x, uv[i].y, x, y);
  xy[i] = make_double2(x, y);
  if (ok) ok[i] = good ? 1 : 0;
}

// One Gauss-Newton system of the pose-only bundle adjustment over the masked corners of this view:
//   r = pi(R X + t) - x_n,  R <- R exp(delta) (right increment), t <- t + dt;  Huber(1.345) on |r| as in ba_options_ (:44-46)
struct PoseSys { double H[6][6]; double g[6]; double cost; };
ICC_D void pose_system(const PoseProblem& Q, const double2* xy, const unsigned char* use, int c0, int c1, double zref, Q4 q, V3 t, PoseSys& S, bool with_jacobian) {
  const int lane = threadIdx.x & 31;
  double H[21], g[6], cost = 0.0;
#pragma unroll
  for (int i = 0; i < 21; ++i) H[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) g[i] = 0.0;
  const double hub = 1.345;
  for (int c = c0 + lane; c < c1; c += 32) {
    if (!use[c]) continue;
    const double4 Xb = Q.board[Q.pid[c]];
    const V3 X = v3(Xb.x / Xb.w, Xb.y / Xb.w, Xb.z / Xb.w - zref);
    const V3 RX = qrot(q, X), Pc = RX + t;
    const double iz = 1.0 / Pc.z, u = Pc.x * iz, v = Pc.y * iz;
    const double r0 = u - xy[c].x, r1 = v - xy[c].y, rn = sqrt(r0 * r0 + r1 * r1);
    // Huber: rho(s) = s for s <= a^2, 2 a sqrt(s) - a^2 beyond; cost = 1/2 rho(|r|^2); IRLS weight = rho'
    const double w = rn <= hub ? 1.0 : hub / rn;
    cost += rn <= hub ? 0.5 * rn * rn : hub * rn - 0.5 * hub * hub;
    if (!with_jacobian || !(Pc.z > 0.0)) { if (!(Pc.z > 0.0)) cost += 1e6; continue; }
    // d pi / d Pc rows, then d Pc / d(delta, t) = [-R [X]x | I]  ->  row_delta = -(row R) x X ... = X x (R^T row)
    const V3 a0 = v3(iz, 0.0, -u * iz), a1 = v3(0.0, iz, -v * iz);
    const V3 b0 = cross(X, qrot_inv(q, a0)), b1 = cross(X, qrot_inv(q, a1));
    const double J0[6] = {b0.x, b0.y, b0.z, a0.x, a0.y, a0.z}, J1[6] = {b1.x, b1.y, b1.z, a1.x, a1.y, a1.z};
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      g[i] += w * (J0[i] * r0 + J1[i] * r1);
#pragma unroll
      for (int j = 0; j <= i; ++j) { H[idx] += w * (J0[i] * J0[j] + J1[i] * J1[j]); ++idx; }
    }
  }
  S.cost = wsum(cost);
  if (!with_jacobian) return;
  int idx = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    S.g[i] = wsum(g[i]);
#pragma unroll
    for (int j = 0; j <= i; ++j) { const double v = wsum(H[idx]); S.H[i][j] = v; S.H[j][i] = v; ++idx; }
  }
}

// Levenberg-Marquardt on the 6 pose parameters (all lanes run the identical scalar logic on warp-reduced sums)
ICC_D void pose_refine(const PoseProblem& Q, const double2* xy, const unsigned char* use, int c0, int c1, double zref, Q4& q, V3& t) {
  PoseSys S;
  pose_system(Q, xy, use, c0, c1, zref, q, t, S, true);
  double lambda = 1e-4, cost = S.cost;
  for (int it = 0; it < 50; ++it) {
    double A[6][6], b[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      b[i] = -S.g[i];
#pragma unroll
      for (int j = 0; j < 6; ++j) A[i][j] = S.H[i][j];
      A[i][i] += lambda * (S.H[i][i] + 1e-12);
    }
    if (!chol_solve<6>(A, b)) { lambda *= 10.0; if (lambda > 1e12) break; continue; }
    const Q4 qn = qnormalized(qmul(q, so3_exp(v3(b[0], b[1], b[2]))));
    const V3 tn = t + v3(b[3], b[4], b[5]);
    PoseSys Sn;
    pose_system(Q, xy, use, c0, c1, zref, qn, tn, Sn, false);
    const double step2 = b[0] * b[0] + b[1] * b[1] + b[2] * b[2] + b[3] * b[3] + b[4] * b[4] + b[5] * b[5];
    if (Sn.cost < cost) {
      const double dec = cost - Sn.cost;
      q = qn; t = tn; cost = Sn.cost; lambda = fmax(lambda * 0.1, 1e-12);
      if (dec <= (Q.lm_rel_tol > 0.0 ? Q.lm_rel_tol : 1e-15) * cost || step2 < 1e-28) break;
      pose_system(Q, xy, use, c0, c1, zref, q, t, S, true);
    } else {
      if (step2 < 1e-28) break;
      lambda *= 10.0;
      if (lambda > 1e12) break;
    }
  }
}

// Normalised DLT homography board plane -> image of one view, by one warp (every lane returns the same values):
// marks the usable corners in use[], returns their count n, the board plane height zref and H (row-major, acting on
// (X, Y, 1) with Z = zref); false = too few corners / degenerate / non-planar target.
ICC_D bool board_homography(const PoseProblem& Q, const double2* __restrict__ xy, const int* __restrict__ okc, unsigned char* __restrict__ use, int c0, int c1,
                            double& n_out, double& zref_out, double (&Hm)[9]) {
  const int lane = threadIdx.x & 31;
  // ---- usable corners, board plane, Hartley normalisation ------------------------------------------------------------------
  double n = 0.0, sX = 0.0, sY = 0.0, sZ = 0.0, sx = 0.0, sy = 0.0;
  for (int c = c0 + lane; c < c1; c += 32) {
    const int id = Q.pid[c];
    const bool u = (okc == nullptr || okc[c] != 0) && id >= 0 && id < Q.n_points;
    use[c] = u ? 1 : 0;
    if (!u) continue;
    const double4 Xb = Q.board[id];
    n += 1.0; sX += Xb.x / Xb.w; sY += Xb.y / Xb.w; sZ += Xb.z / Xb.w; sx += xy[c].x; sy += xy[c].y;
  }
  n = wsum(n); sX = wsum(sX); sY = wsum(sY); sZ = wsum(sZ); sx = wsum(sx); sy = wsum(sy);
  __syncwarp();
  n_out = n;
  if (c1 - c0 < Q.min_points || n < 6.0) return false;   // pose_estimator.cc:140 (all corners counted), :65
  const double mX = sX / n, mY = sY / n, zref = sZ / n, mx = sx / n, my = sy / n;
  zref_out = zref;
  double dB = 0.0, dI = 0.0, dz = 0.0;
  for (int c = c0 + lane; c < c1; c += 32) {
    if (!use[c]) continue;
    const double4 Xb = Q.board[Q.pid[c]];
    const double X = Xb.x / Xb.w - mX, Y = Xb.y / Xb.w - mY, x = xy[c].x - mx, y = xy[c].y - my;
    dB += sqrt(X * X + Y * Y); dI += sqrt(x * x + y * y); dz = fmax(dz, fabs(Xb.z / Xb.w - zref));
  }
  dB = wsum(dB) / n; dI = wsum(dI) / n;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dz = fmax(dz, __shfl_xor_sync(0xffffffffu, dz, o));
  if (!(dB > 0.0) || !(dI > 0.0) || dz > 0.05 * dB) return false;   // degenerate, or too far from a plane for the homography to start the refinement
                                                                     // (a few per cent of the board size -- refined board points of a printed target -- are fine: pose_system uses the real 3-D points)
  const double sB = sqrt(2.0) / dB, sI = sqrt(2.0) / dI;
  // ---- homography (h33 = 1) from the normal equations of the DLT rows -------------------------------------------------------
  double M[36], v8[8];
#pragma unroll
  for (int i = 0; i < 36; ++i) M[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) v8[i] = 0.0;
  for (int c = c0 + lane; c < c1; c += 32) {
    if (!use[c]) continue;
    const double4 Xb = Q.board[Q.pid[c]];
    const double X = (Xb.x / Xb.w - mX) * sB, Y = (Xb.y / Xb.w - mY) * sB, x = (xy[c].x - mx) * sI, y = (xy[c].y - my) * sI;
    const double ra[8] = {X, Y, 1.0, 0.0, 0.0, 0.0, -x * X, -x * Y}, rb[8] = {0.0, 0.0, 0.0, X, Y, 1.0, -y * X, -y * Y};
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v8[i] += ra[i] * x + rb[i] * y;
#pragma unroll
      for (int j = 0; j <= i; ++j) { M[idx] += ra[i] * ra[j] + rb[i] * rb[j]; ++idx; }
    }
  }
  double A8[8][8], h[8];
  {
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      h[i] = wsum(v8[i]);
#pragma unroll
      for (int j = 0; j <= i; ++j) { const double v = wsum(M[idx]); A8[i][j] = v; A8[j][i] = v; ++idx; }
    }
  }
  if (!chol_solve<8>(A8, h)) return false;
  // de-normalise: H = T_img^-1 Hn T_board,  T_board = [sB 0 -sB mX; 0 sB -sB mY; 0 0 1],  T_img^-1 = [1/sI 0 mx; 0 1/sI my; 0 0 1]
  double Hn[9] = {h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], 1.0}, G[9];
#pragma unroll
  for (int r = 0; r < 3; ++r) { G[3 * r] = Hn[3 * r] * sB; G[3 * r + 1] = Hn[3 * r + 1] * sB; G[3 * r + 2] = Hn[3 * r + 2] - sB * (Hn[3 * r] * mX + Hn[3 * r + 1] * mY); }
#pragma unroll
  for (int cidx = 0; cidx < 3; ++cidx) { Hm[cidx] = G[cidx] / sI + mx * G[6 + cidx]; Hm[3 + cidx] = G[3 + cidx] / sI + my * G[6 + cidx]; Hm[6 + cidx] = G[6 + cidx]; }
  return true;
}

__global__ void __launch_bounds__(128) board_pose_kernel(PoseProblem Q, const double2* __restrict__ xy, const int* __restrict__ okc, unsigned char* __restrict__ use,
                                                        double* q_out, double* p_out, double* __restrict__ err_out, int* valid_out) {
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (f >= Q.n_frames) return;
  const int c0 = Q.f_off[f], c1 = Q.f_off[f + 1];
  auto fail = [&]() {
    if (lane == 0) { q_out[4 * f] = 0.0; q_out[4 * f + 1] = 0.0; q_out[4 * f + 2] = 0.0; q_out[4 * f + 3] = 1.0; p_out[3 * f] = 0.0; p_out[3 * f + 1] = 0.0; p_out[3 * f + 2] = 0.0; err_out[f] = 0.0; valid_out[f] = 0; }
  };
  double n = 0.0, zref = 0.0;
  Q4 q; V3 t;
  if (Q.refine_only) {
    // PoseEstimator::OptimizeAllPoses (pose_estimator.cc:226-236): BundleAdjustView again from the stored pose -- no homography, so the
    // board may have left its plane (optimised board points); views that were dropped stay dropped and keep their outputs
    if (!valid_out[f]) return;
    for (int c = c0 + lane; c < c1; c += 32) {
      const int id = Q.pid[c];
      const bool u = okc[c] != 0 && id >= 0 && id < Q.n_points;
      use[c] = u ? 1 : 0;
      if (u) n += 1.0;
    }
    n = wsum(n);
    __syncwarp();
    if (c1 - c0 < Q.min_points || n < 6.0) { fail(); return; }
    const Q4 qw = qnormalized(q4(q_out[4 * f], q_out[4 * f + 1], q_out[4 * f + 2], q_out[4 * f + 3]));
    q = qconj(qw);
    t = -qrot(q, v3(p_out[3 * f], p_out[3 * f + 1], p_out[3 * f + 2]));
  } else {
    double Hm[9];
    if (!board_homography(Q, xy, okc, use, c0, c1, n, zref, Hm)) { fail(); return; }
    // H ~ [r1 r2 t] (points on the plane z = zref, shifted to z = 0): scale, cheirality, Gram-Schmidt
    V3 h1 = v3(Hm[0], Hm[3], Hm[6]), h2 = v3(Hm[1], Hm[4], Hm[7]), h3 = v3(Hm[2], Hm[5], Hm[8]);
    double s = 2.0 / (sqrt(dot(h1, h1)) + sqrt(dot(h2, h2)));
    if (h3.z * s < 0.0) s = -s;
    V3 r1 = s * h1, r2 = s * h2;
    t = s * h3;
    r1 = (1.0 / sqrt(dot(r1, r1))) * r1;
    r2 = r2 - dot(r1, r2) * r1; r2 = (1.0 / sqrt(dot(r2, r2))) * r2;
    const V3 r3 = cross(r1, r2);
    q = quat_from_columns(r1, r2, r3);          // R_cw
  }
  // ---- refinement, inlier selection with the reference's squared normalised threshold, refinement on the inliers -------------
  pose_refine(Q, xy, use, c0, c1, zref, q, t);
  double n_in = 0.0, n_out = 0.0;
  for (int c = c0 + lane; c < c1; c += 32) {
    if (!use[c]) continue;
    const double4 Xb = Q.board[Q.pid[c]];
    const V3 Pc = qrot(q, v3(Xb.x / Xb.w, Xb.y / Xb.w, Xb.z / Xb.w - zref)) + t;
    const double r0 = Pc.x / Pc.z - xy[c].x, r1e = Pc.y / Pc.z - xy[c].y;
    const bool in = Pc.z > 0.0 && r0 * r0 + r1e * r1e < Q.thresh_sq;
    if (in) n_in += 1.0; else { n_out += 1.0; use[c] = 0; }
  }
  n_in = wsum(n_in); n_out = wsum(n_out);
  __syncwarp();
  if (n_in < 6.0) { fail(); return; }
  if (n_out > 0.0) pose_refine(Q, xy, use, c0, c1, zref, q, t);
  // ---- mean reprojection error of the kept observations (:164-181; the view camera is the unit pinhole) ----------------------
  double e = 0.0;
  for (int c = c0 + lane; c < c1; c += 32) {
    if (!use[c]) continue;
    const double4 Xb = Q.board[Q.pid[c]];
    const V3 Pc = qrot(q, v3(Xb.x / Xb.w, Xb.y / Xb.w, Xb.z / Xb.w - zref)) + t;
    const double r0 = Pc.x / Pc.z - xy[c].x, r1e = Pc.y / Pc.z - xy[c].y;
    e += sqrt(r0 * r0 + r1e * r1e);
  }
  e = wsum(e) / n_in;
  if (lane == 0) {
    // undo the plane shift: Pc = R (X - zref e_z) + t  =>  t_full = t - zref R e_z ; camera centre p_wc = -R^T t_full ; q_wc = q^*
    const V3 tf = t - zref * qrot(q, v3(0.0, 0.0, 1.0));
    const V3 pw = -qrot_inv(q, tf);
    const Q4 qw = qconj(q);
    const bool keep = e <= Q.max_err;
    q_out[4 * f] = qw.x; q_out[4 * f + 1] = qw.y; q_out[4 * f + 2] = qw.z; q_out[4 * f + 3] = qw.w;
    p_out[3 * f] = pw.x; p_out[3 * f + 1] = pw.y; p_out[3 * f + 2] = pw.z;
    err_out[f] = e; valid_out[f] = keep ? 1 : 0;
  }
}


// Focal length of a view from its board homography on principal-point-centred pixels (square pixels, zero skew):
// H ~ K [r1 r2 t], K = diag(f, f, 1), r1 . r2 = 0 and |r1| = |r2| give two linear equations a_i + f^2 b_i = 0 (Zhang's
// constraints); least squares over both.  f2_out[v] = f^2, or 0 when the view is unusable / the estimate is not positive.
__global__ void __launch_bounds__(128) board_focal_kernel(PoseProblem Q, const double2* __restrict__ uv, double cx, double cy, double2* __restrict__ xy,
                                                         unsigned char* __restrict__ use, double* __restrict__ f2_out) {
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (f >= Q.n_frames) return;
  const int c0 = Q.f_off[f], c1 = Q.f_off[f + 1];
  for (int c = c0 + lane; c < c1; c += 32) xy[c] = make_double2(uv[c].x - cx, uv[c].y - cy);
  __syncwarp();
  double n = 0.0, zref = 0.0, H[9], f2 = 0.0;
  if (board_homography(Q, xy, nullptr, use, c0, c1, n, zref, H)) {
    const double a1 = H[0] * H[1] + H[3] * H[4], b1 = H[6] * H[7];
    const double a2 = H[0] * H[0] + H[3] * H[3] - H[1] * H[1] - H[4] * H[4], b2 = H[6] * H[6] - H[7] * H[7];
    const double den = b1 * b1 + b2 * b2;
    if (den > 0.0) f2 = -(a1 * b1 + a2 * b2) / den;
    if (!(f2 > 0.0) || !isfinite(f2)) f2 = 0.0;
  }
  if (lane == 0) f2_out[f] = f2;
}

__global__ void pinhole_normalize_kernel(int n, const double2* __restrict__ uv, double cx, double cy, double inv_f, double2* __restrict__ xy, int* __restrict__ ok) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  xy[i] = make_double2((uv[i].x - cx) * inv_f, (uv[i].y - cy) * inv_f);
  ok[i] = 1;
}


// theia::BundleAdjustTracks with every camera constant (PoseEstimator::OptimizeBoardPoints, pose_estimator.cc:193-224; the tail of
// CameraCalibrator::RunCalibration, camera_calibrator.cc:207-216): with the cameras fixed every board point is its own 3-parameter
// problem, so ONE WARP owns one point: lanes stride over the point's observations (CSR by point, built on the host), the 3x3
// Levenberg-Marquardt system is warp-reduced and solved in registers by every lane.  Residual: Huber(1.345) on
//   normalised == 1 :  (R_cw (X - c)).xy / z - x_n          (the pose estimator's unit-pinhole views with undistorted features)
//   normalised == 0 :  CameraToPixelCoordinates(intr, R_cw (X - c)) - pixel     (the camera calibrator's views)
struct PointSys { double H[3][3]; double g[3]; double cost; };
ICC_D void point_system(const PointProblem& Q, int o0, int o1, V3 X, PointSys& S, bool with_jacobian) {
  const int lane = threadIdx.x & 31;
  double H[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0}, cost = 0.0;
  const double hub = Q.huber;
  for (int o = o0 + lane; o < o1; o += 32) {
    const int c = Q.pt_obs[o], v = Q.obs_view[c];
    const Q4 q = q4(Q.q_cw[4 * v], Q.q_cw[4 * v + 1], Q.q_cw[4 * v + 2], Q.q_cw[4 * v + 3]);
    const V3 pc = qrot(q, X - v3(Q.cam_c[3 * v], Q.cam_c[3 * v + 1], Q.cam_c[3 * v + 2]));
    double r0, r1; V3 a0, a1;
    if (Q.normalized) {
      if (!(pc.z > 0.0)) { cost += 1e6; continue; }
      const double iz = 1.0 / pc.z, u = pc.x * iz, w = pc.y * iz;
      r0 = u - Q.meas[c].x; r1 = w - Q.meas[c].y;
      a0 = v3(iz, 0.0, -u * iz); a1 = v3(0.0, iz, -w * iz);
    } else {
      const Proj pr = project(Q.model, Q.intr, pc, true);
      if (!pr.ok) { cost += 1e10; continue; }
      r0 = pr.u - Q.meas[c].x; r1 = pr.v - Q.meas[c].y;
      a0 = v3(pr.J[0], pr.J[1], pr.J[2]); a1 = v3(pr.J[3], pr.J[4], pr.J[5]);
    }
    const double rn = sqrt(r0 * r0 + r1 * r1), wgt = rn <= hub ? 1.0 : hub / rn;
    cost += rn <= hub ? 0.5 * rn * rn : hub * rn - 0.5 * hub * hub;
    if (!with_jacobian) continue;
    const V3 j0 = qrot_inv(q, a0), j1 = qrot_inv(q, a1);   // rows of J_proj R_cw
    g[0] += wgt * (j0.x * r0 + j1.x * r1); g[1] += wgt * (j0.y * r0 + j1.y * r1); g[2] += wgt * (j0.z * r0 + j1.z * r1);
    H[0] += wgt * (j0.x * j0.x + j1.x * j1.x); H[1] += wgt * (j0.x * j0.y + j1.x * j1.y); H[2] += wgt * (j0.x * j0.z + j1.x * j1.z);
    H[3] += wgt * (j0.y * j0.y + j1.y * j1.y); H[4] += wgt * (j0.y * j0.z + j1.y * j1.z); H[5] += wgt * (j0.z * j0.z + j1.z * j1.z);
  }
  S.cost = wsum(cost);
  if (!with_jacobian) return;
#pragma unroll
  for (int i = 0; i < 3; ++i) S.g[i] = wsum(g[i]);
  const double h0 = wsum(H[0]), h1 = wsum(H[1]), h2 = wsum(H[2]), h3 = wsum(H[3]), h4 = wsum(H[4]), h5 = wsum(H[5]);
  S.H[0][0] = h0; S.H[0][1] = h1; S.H[0][2] = h2; S.H[1][0] = h1; S.H[1][1] = h3; S.H[1][2] = h4; S.H[2][0] = h2; S.H[2][1] = h4; S.H[2][2] = h5;
}

__global__ void __launch_bounds__(128) point_refine_kernel(PointProblem Q, double4* __restrict__ board_out, int* __restrict__ optimized) {
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (p >= Q.n_points) return;
  const int o0 = Q.pt_off[p], o1 = Q.pt_off[p + 1];
  const double4 Xb = Q.board_in[p];
  if (o1 - o0 <= Q.min_obs) { if (lane == 0) { board_out[p] = Xb; optimized[p] = 0; } return; }   // pose_estimator.cc:202-206
  V3 X = v3(Xb.x / Xb.w, Xb.y / Xb.w, Xb.z / Xb.w);
  PointSys S;
  point_system(Q, o0, o1, X, S, true);
  double lambda = 1e-4, cost = S.cost;
  for (int it = 0; it < 50; ++it) {
    double A[3][3], b[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      b[i] = -S.g[i];
#pragma unroll
      for (int j = 0; j < 3; ++j) A[i][j] = S.H[i][j];
      A[i][i] += lambda * (S.H[i][i] + 1e-12);
    }
    if (!chol_solve<3>(A, b)) { lambda *= 10.0; if (lambda > 1e12) break; continue; }
    const V3 Xn = X + v3(b[0], b[1], b[2]);
    PointSys Sn;
    point_system(Q, o0, o1, Xn, Sn, false);
    const double step2 = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
    if (Sn.cost < cost) {
      const double dec = cost - Sn.cost;
      X = Xn; cost = Sn.cost; lambda = fmax(lambda * 0.1, 1e-12);
      if (dec <= 1e-15 * cost || step2 < 1e-30) break;
      point_system(Q, o0, o1, X, S, true);
    } else {
      if (step2 < 1e-30) break;
      lambda *= 10.0;
      if (lambda > 1e12) break;
    }
  }
  if (lane == 0) { board_out[p] = make_double4(X.x, X.y, X.z, 1.0); optimized[p] = 1; }
}

}  // namespace

void launch_unproject(int model, const double* intr10, int n, const double2* uv, double2* xy, int* ok, cudaStream_t st) {
  if (n <= 0) return;
  Intr K; for (int i = 0; i < 10; ++i) K.k[i] = intr10[i];
  unproject_kernel<<<(n + 127) / 128, 128, 0, st>>>(model, K, n, uv, xy, ok);
  count_launch();
}

void launch_board_poses(const PoseProblem& Q, const double2* xy, const int* ok, unsigned char* use, double* q_wc, double* p_wc, double* err, int* valid, cudaStream_t st) {
  if (Q.n_frames <= 0) return;
  board_pose_kernel<<<(Q.n_frames + 3) / 4, 128, 0, st>>>(Q, xy, ok, use, q_wc, p_wc, err, valid);
  count_launch();
}

void launch_board_focal(const PoseProblem& Q, const double2* uv, double cx, double cy, double2* xy, unsigned char* use, double* f2, cudaStream_t st) {
  if (Q.n_frames <= 0) return;
  board_focal_kernel<<<(Q.n_frames + 3) / 4, 128, 0, st>>>(Q, uv, cx, cy, xy, use, f2);
  count_launch();
}

void launch_pinhole_normalize(int n, const double2* uv, double cx, double cy, double f, double2* xy, int* ok, cudaStream_t st) {
  if (n <= 0) return;
  pinhole_normalize_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, uv, cx, cy, 1.0 / f, xy, ok);
  count_launch();
}

void launch_point_refine(const PointProblem& Q, double4* board_out, int* optimized, cudaStream_t st) {
  if (Q.n_points <= 0) return;
  point_refine_kernel<<<(Q.n_points + 3) / 4, 128, 0, st>>>(Q, board_out, optimized);
  count_launch();
}

}  // namespace icc
