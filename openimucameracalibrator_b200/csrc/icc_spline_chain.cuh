// The per-observation SO(3) spline chain and its analytic knot Jacobian rows, shared by the vision and IMU kernels of icc_eval.cu.
//
// Kept in a header (and __host__ __device__) so that tests/test_host_device_math.py can compile exactly this code for the CPU and
// check the Jacobian rows against finite differences of the spline rotation; the kernels inline it unchanged.
// Recipe: header of icc_device_math.cuh; reference counterpart CeresSplineHelper::evaluate_lie (basalt_spline/ceres_spline_helper.h:101-187)
// under Ceres autodiff + LieLocalParameterization (right increments on the knots).
#pragma once
#include "icc_device_math.cuh"

namespace icc {

constexpr int LDJ = 36;   // rows per tile column (32 + 4 pad: stride = 4 mod 16 => conflict-free DMMA fragment loads)

struct WarpCtx {
  Q4 q[6];
  V3 d[5];
  V3 dh[5];          // unit axes d_i / |d_i| (0 when the increment vanishes)
  double th[5], ith[5];   // |d_i| and its reciprocal (0 when the increment vanishes): no division, sqrt or norm per observation
  M3 jri[5];
  V3 p[6];
  V3 ba[3], bg[3];
  int gidx[64];
  int gidx2[40];
};

// Lane `i` < 5 of a work item: log increment d_i = log(R_i^-1 R_{i+1}) of the staged window, its norm / unit axis and Jr^-1(d_i)
template <bool JAC>
ICC_HD void stage_so3_increment(WarpCtx* wc, int i) {
  const V3 d = so3_log(qmul(qconj(wc->q[i]), wc->q[i + 1]));
  wc->d[i] = d;
  const double th = sqrt(dot(d, d)), ith = th > 1e-150 ? 1.0 / th : 0.0;
  wc->th[i] = th; wc->ith[i] = ith; wc->dh[i] = ith * d;
  if (JAC) wc->jri[i] = so3_jr_inv(d);
}

// lambda * (x Jr(lambda d)) for a row vector x, with the per-item unit axis dh and the two per-observation coefficients of Chain
ICC_HD V3 lam_row_jr(V3 x, V3 dh, double lam, double c1, double c2) {
  const V3 xd = cross(x, dh);
  return lam * x - c1 * xd + c2 * cross(xd, dh);
}

// Per-observation spline rotation chain shared by all residual types.
struct Chain {
  Q4 A[5];          // exp(lambda_i d_i), i = 1..5
  double c1[5], c2[5];   // lambda * row Jr(lambda d) = lambda x - c1 (x x dh) + c2 ((x x dh) x dh):  c1 = (1 - cos phi)/theta, c2 = lambda - sin(phi)/theta
  double lam[5], dlam[5];
  Q4 q;             // R_w_i
};
ICC_HD void build_chain(const WarpCtx* wc, double u, Chain& ch) {
  cum_coeffs6(u, ch.lam, ch.dlam);
  Q4 q = wc->q[0];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    // exp(lambda d) about the fixed axis dh: one sincos of the half angle; phi = lambda theta
    double sn, cs;
    sincos(0.5 * ch.lam[i] * wc->th[i], &sn, &cs);
    const V3 ax = wc->dh[i];
    const Q4 eq = q4(sn * ax.x, sn * ax.y, sn * ax.z, cs);
    ch.A[i] = eq; ch.c1[i] = 2.0 * sn * sn * wc->ith[i]; ch.c2[i] = ch.lam[i] - 2.0 * sn * cs * wc->ith[i];
    q = qmul(q, eq);
  }
  ch.q = q;
}

// Given the row covector m_theta = d r / d theta (right increment of R_w_i), write d r / d eps_j for the six SO(3) knots
// into tile columns [0, 18) of row `lane`, and return sum_i dlam_i <w_i, d_i>  (= d r / d u through the rotation).
ICC_HD double so3_knot_row(const WarpCtx* wc, const Chain& ch, V3 m_theta, double* __restrict__ Jt, int lane, double scale) {
  V3 w = m_theta;                  // w_5
  V3 z_next = v3(0, 0, 0);         // Jr^-1_{j+1} z_{j+1} handled below
  double du = 0.0;
#pragma unroll
  for (int i = 4; i >= 0; --i) {   // knot pair (i, i+1): increment index i+1 in the text, array index i
    const V3 di = wc->d[i];
    du += ch.dlam[i] * dot(w, di);
    const V3 z = lam_row_jr(w, wc->dh[i], ch.lam[i], ch.c1[i], ch.c2[i]);
    // knot i+1 receives  z Jr^-1(d_i)  (row-vector times matrix) minus the contribution found in the previous iteration
    const V3 up = mulT(wc->jri[i], z) - z_next;
    Jt[(3 * (i + 1) + 0) * LDJ + lane] = scale * up.x;
    Jt[(3 * (i + 1) + 1) * LDJ + lane] = scale * up.y;
    Jt[(3 * (i + 1) + 2) * LDJ + lane] = scale * up.z;
    z_next = mul(wc->jri[i], z);   // z Jl^-1(d_i) = Jr^-1(d_i) z, subtracted from knot i
    w = qrot(ch.A[i], w);          // w_{i}  (P_i^T applied)
  }
  const V3 k0 = w - z_next;
  Jt[0 * LDJ + lane] = scale * k0.x; Jt[1 * LDJ + lane] = scale * k0.y; Jt[2 * LDJ + lane] = scale * k0.z;
  return du;
}

}  // namespace icc
