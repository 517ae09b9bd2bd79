// Damped normal-equation solve + manifold update kernels (sm_100a).
//
// Replaces what ceres::Solve does between two Jacobian evaluations for the reference's spline problem
// (SplineTrajectoryEstimator::Optimize, include/OpenCameraCalibrator/core/spline_trajectory_estimator.impl.h:254-276):
//   * Jacobi column scaling + Levenberg-Marquardt diagonal  (Ceres TrustRegionMinimizer / LevenbergMarquardtStrategy)
//   * SPARSE_NORMAL_CHOLESKY on (J^T J + D^T D)              -> exact LDL^T of the banded (spline knots, time ordered)
//     + bordered (T_i_c, gravity, line delay, bias knots) system, i.e. the spline control points are eliminated first and
//     the small dense Schur complement of the border is factored last
//   * LieLocalParameterization::Plus on every SO(3) knot and on T_i_c (basalt_spline/ceres_local_param.h:84-92)
// The factorisation is a sequential recurrence along time; v1 runs it in ONE thread block with the active window of the
// band resident in shared memory (one __syncthreads per eliminated column, no square roots: LDL^T).
#include "icc_device_math.cuh"
#include "icc_kernels.h"

namespace icc {

void count_launch();

namespace {

constexpr int SOLVE_THREADS = 512;

__global__ void scale_kernel(DeviceProblem P, double* scale, int jacobi, double* scal) {
  const int n = P.nk + P.nb;
  __shared__ double red[32];
  double gmax = 0.0;
  const double* g = P.ne + P.ne_off_g;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double d = i < P.nk ? P.ne[(int64_t)i * P.ldb] : P.ne[P.ne_off_C + (int64_t)(i - P.nk) * P.nb + (i - P.nk)];
    if (scale) scale[i] = jacobi ? 1.0 / (1.0 + sqrt(d)) : 1.0;
    gmax = fmax(gmax, fabs(g[i]));
  }
  for (int o = 16; o > 0; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = gmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) m = fmax(m, red[w]);
    // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(scal + SC_GRAD_MAX), (unsigned long long)__double_as_longlong(m));
  }
}

// Workspace layout (doubles): Lb[nk*ldb] | Le[nk*nbp] | d2[n] | y[n] | t[nk]
struct SolveWs { double* Lb; double* Le; double* d2; double* y; double* t; };
__host__ __device__ inline SolveWs carve(double* ws, int nk, int nb, int ldb) {
  SolveWs w; const int nbp = nb + 1;
  w.Lb = ws; w.Le = w.Lb + (int64_t)nk * ldb; w.d2 = w.Le + (int64_t)nk * nbp; w.y = w.d2 + (nk + nb); w.t = w.y + (nk + nb);
  return w;
}

__device__ __forceinline__ int tri_row(int idx) {   // idx = r(r+1)/2 + c, 0 <= c <= r  ->  r
  int r = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
  while ((r + 1) * (r + 2) / 2 <= idx) ++r;
  while (r * (r + 1) / 2 > idx) --r;
  return r;
}

__global__ void __launch_bounds__(SOLVE_THREADS) solve_kernel(DeviceProblem P, const double* __restrict__ scale, SolveParams sp, double* wsp, double* delta,
                                                               double* scal, int PB) {
  extern __shared__ __align__(16) double sm[];
  const int nk = P.nk, nb = P.nb, kd = P.kd, ldb = P.ldb, nbp = nb + 1, n = nk + nb;
  const int CL = ldb + nbp;            // column length inside the window: band part + border part + rhs
  const int WS = PB + kd;              // window columns
  double* W = sm;                                     // WS * CL
  double* Cs = W + (int64_t)WS * CL;                  // nbp * nbp   (lower; row nb = rhs)
  double* xb = Cs + nbp * nbp;                        // nbp
  unsigned short* pr = reinterpret_cast<unsigned short*>(xb + nbp + 1);   // pair table for the band triangle: (r << 8) | c
  const int np_full = kd * (kd + 1) / 2;
  __shared__ int s_ok;
  const int tid = threadIdx.x, nt = blockDim.x;
  const double* band = P.ne; const double* E = P.ne + P.ne_off_E; const double* C = P.ne + P.ne_off_C; const double* g = P.ne + P.ne_off_g;
  SolveWs ws = carve(wsp, nk, nb, ldb);

  if (tid == 0) s_ok = 1;
  for (int idx = tid; idx < np_full; idx += nt) { const int r0 = tri_row(idx); const int c0 = idx - r0 * (r0 + 1) / 2; pr[idx] = (unsigned short)(((r0 + 1) << 8) | (c0 + 1)); }   // 1 <= c <= r <= kd
  // LM diagonal D^2 = clamp(diag(S H S)) / radius   (LevenbergMarquardtStrategy::ComputeStep)
  for (int i = tid; i < n; i += nt) {
    const double d = i < nk ? band[(int64_t)i * ldb] : C[(int64_t)(i - nk) * nb + (i - nk)];
    const double s = scale[i];
    ws.d2[i] = fmin(fmax(d * s * s, sp.min_diag), sp.max_diag) / sp.radius;
  }
  __syncthreads();
  for (int idx = tid; idx < nbp * nbp; idx += nt) {
    const int b = idx / nbp, c = idx % nbp;
    double v = 0.0;
    if (c <= b) {
      if (b < nb) { v = C[(int64_t)b * nb + c] * scale[nk + b] * scale[nk + c]; if (b == c) v += ws.d2[nk + b]; }
      else if (c < nb) v = -g[nk + c] * scale[nk + c];
    }
    Cs[idx] = v;
  }
  __syncthreads();

  // ---------------- banded part: LDL^T column by column, window resident in shared memory -----------------------
  bool ok = true;
  for (int j0 = 0; j0 < nk && ok; j0 += PB) {
    const int first = j0 == 0 ? 0 : j0 + kd, last = min(nk, j0 + PB + kd);
    for (int idx = tid; idx < (last - first) * CL; idx += nt) {
      const int col = first + idx / CL, e = idx % CL;
      double v;
      if (e < ldb) { const int i = col + e; v = i < nk ? band[(int64_t)col * ldb + e] * scale[col] * scale[i] : 0.0; if (e == 0) v += ws.d2[col]; }
      else { const int b = e - ldb; v = b < nb ? E[(int64_t)col * nb + b] * scale[col] * scale[nk + b] : -g[col] * scale[col]; }
      W[(int64_t)(col % WS) * CL + e] = v;
    }
    __syncthreads();
    const int jend = min(j0 + PB, nk);
    for (int j = j0; j < jend; ++j) {
      const double* cj = W + (int64_t)(j % WS) * CL;
      const double piv = cj[0];
      if (!(piv > 0.0) || !isfinite(piv)) { ok = false; break; }
      const double inv = 1.0 / piv;
      const int m = min(kd, nk - 1 - j);
      const int np = m * (m + 1) / 2, nbb = nbp * m, ncc = nbp * (nbp + 1) / 2;
      for (int idx = tid; idx < np + nbb + ncc; idx += nt) {
        if (idx < np) {
          const int r = pr[idx] >> 8, c = pr[idx] & 255;
          W[(int64_t)((j + c) % WS) * CL + (r - c)] -= cj[r] * cj[c] * inv;
        } else if (idx < np + nbb) {
          const int k = idx - np; const int b = k / m, c = k % m + 1;
          W[(int64_t)((j + c) % WS) * CL + ldb + b] -= cj[ldb + b] * cj[c] * inv;
        } else {
          const int k = idx - np - nbb; const int b = tri_row(k), c = k - b * (b + 1) / 2;
          Cs[b * nbp + c] -= cj[ldb + b] * cj[ldb + c] * inv;
        }
      }
      __syncthreads();
    }
    if (!ok) break;
    for (int idx = tid; idx < (jend - j0) * CL; idx += nt) {
      const int col = j0 + idx / CL, e = idx % CL;
      const double v = W[(int64_t)(col % WS) * CL + e];
      if (e < ldb) ws.Lb[(int64_t)col * ldb + e] = v; else ws.Le[(int64_t)col * nbp + (e - ldb)] = v;
    }
    __syncthreads();
  }
  // ---------------- border: dense LDL^T of the Schur complement, rhs carried as the last row --------------------
  for (int j = 0; j < nb && ok; ++j) {
    const double piv = Cs[j * nbp + j];
    if (!(piv > 0.0) || !isfinite(piv)) { ok = false; break; }
    const double inv = 1.0 / piv;
    const int mrem = nbp - 1 - j;                 // rows j+1..nb (incl. rhs row)
    for (int idx = tid; idx < mrem * (mrem + 1) / 2; idx += nt) {
      const int r = tri_row(idx), c = idx - r * (r + 1) / 2;
      const int b = j + 1 + r, cc = j + 1 + c;
      Cs[b * nbp + cc] -= Cs[b * nbp + j] * Cs[cc * nbp + j] * inv;
    }
    __syncthreads();
  }
  if (!ok) { if (tid == 0) { scal[SC_OK] = 0.0; scal[SC_MODEL_CHANGE] = 0.0; } return; }
  // border back-substitution (warp 0): x_b = (rhs_b - sum_{i>b} L_ib D_b x_i) / D_b with unscaled columns
  if (tid < 32) {
    for (int b = tid; b < nb; b += 32) xb[b] = Cs[nb * nbp + b];
    __syncwarp();
    for (int j = nb - 1; j >= 0; --j) {
      const double xj = xb[j] / Cs[j * nbp + j];
      __syncwarp();
      if (tid == 0) xb[j] = xj;
      for (int i = tid; i < j; i += 32) xb[i] -= Cs[j * nbp + i] * xj;
      __syncwarp();
    }
  }
  __syncthreads();
  for (int b = tid; b < nb; b += nt) ws.y[nk + b] = xb[b];
  // knot right-hand side after removing the border: t_j = rhs_j - sum_b Le[j][b] x_b
  for (int j = tid; j < nk; j += nt) {
    const double* le = ws.Le + (int64_t)j * nbp;
    double t = le[nb];
    for (int b = 0; b < nb; ++b) t -= le[b] * xb[b];
    ws.t[j] = t;
  }
  __syncthreads();
  // ---------------- knot back-substitution, descending, scatter form, one warp drives the recurrence -----------
  // window: columns [lo, hi) with their band entries and running t; x_j = t_j / D_j ; t_{j-r} -= L_{j,j-r} D x_j = W[j-r][r] x_j
  {
    double* Bw = W;                          // (PB + kd) * ldb
    double* tw = W + (int64_t)(PB + kd) * ldb;   // PB + kd
    for (int hi = nk; hi > 0; hi -= PB) {
      const int lo_own = max(0, hi - PB), lo = max(0, lo_own - kd);
      for (int idx = tid; idx < (hi - lo) * ldb; idx += nt) Bw[idx] = ws.Lb[(int64_t)lo * ldb + idx];
      for (int idx = tid; idx < hi - lo; idx += nt) tw[idx] = ws.t[lo + idx];
      __syncthreads();
      if (tid < 32) {
        for (int j = hi - 1; j >= lo_own; --j) {
          const int jl = j - lo;
          const double xj = tw[jl] / Bw[(int64_t)jl * ldb];
          __syncwarp();
          if (tid == 0) tw[jl] = xj;
          for (int r = tid + 1; r <= kd && r <= jl; r += 32) tw[jl - r] -= Bw[(int64_t)(jl - r) * ldb + r] * xj;
          __syncwarp();
        }
      }
      __syncthreads();
      for (int idx = tid; idx < hi - lo_own; idx += nt) ws.y[lo_own + idx] = tw[lo_own - lo + idx];
      for (int idx = tid; idx < lo_own - lo; idx += nt) ws.t[lo + idx] = tw[idx];
      __syncthreads();
    }
  }
  // ---------------- step in the unscaled space, model cost change = 1/2 (y^T D y + y^T rhs) ----------------------
  double part = 0.0;
  for (int i = tid; i < n; i += nt) {
    const double yi = ws.y[i], s = scale[i];
    delta[i] = yi * s;
    part += ws.d2[i] * yi * yi - yi * g[i] * s;
  }
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  double* red = Cs;   // border storage no longer needed
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) { double s = 0.0; for (int w = 0; w < nt / 32; ++w) s += red[w]; scal[SC_MODEL_CHANGE] = 0.5 * s; scal[SC_OK] = 1.0; }
}

// SE3::exp (sophus/se3.hpp:761-783), tangent = (upsilon, omega)
__device__ void se3_exp_dev(const double* a, Q4& q, V3& t) {
  const V3 ups = v3(a[0], a[1], a[2]), om = v3(a[3], a[4], a[5]);
  const double th2 = dot(om, om);
  const ExpOut e = so3_exp_jr(om);
  q = e.q;
  if (th2 < kEps * kEps) { t = qrot(q, ups); return; }   // V = so3.matrix() branch
  // V = I + (1-cos t)/t^2 [om]x + (t - sin t)/t^3 [om]x^2   (the Jr coefficients with the opposite sign on the first term)
  const V3 c1 = cross(om, ups);
  t = ups + e.a * c1 + e.b * cross(om, c1);
}

__global__ void update_kernel(DeviceProblem P, DeviceState cur, DeviceState cand, const double* __restrict__ delta, double max_ba, double max_bg, double* scal) {
  const int total = P.n_so3 + P.n_r3 + P.n_ba + P.n_bg + 1;
  double step = 0.0, xsq = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < P.n_so3) {
      const double4 k = cur.so3[i];
      const int c = P.so3_col[i];
      double4 o = k;
      if (c >= 0) {
        const Q4 r = qnormalized(qmul(q4(k.x, k.y, k.z, k.w), so3_exp(v3(delta[c], delta[c + 1], delta[c + 2]))));
        o = make_double4(r.x, r.y, r.z, r.w);
        step += (o.x - k.x) * (o.x - k.x) + (o.y - k.y) * (o.y - k.y) + (o.z - k.z) * (o.z - k.z) + (o.w - k.w) * (o.w - k.w);
        xsq += k.x * k.x + k.y * k.y + k.z * k.z + k.w * k.w;
      }
      cand.so3[i] = o;
    } else if (i < P.n_so3 + P.n_r3) {
      const int j = i - P.n_so3;
      const double4 k = cur.r3[j];
      const int c = P.r3_col[j];
      double4 o = k;
      if (c >= 0) {
        o = make_double4(k.x + delta[c], k.y + delta[c + 1], k.z + delta[c + 2], 0.0);
        step += (o.x - k.x) * (o.x - k.x) + (o.y - k.y) * (o.y - k.y) + (o.z - k.z) * (o.z - k.z);
        xsq += k.x * k.x + k.y * k.y + k.z * k.z;
      }
      cand.r3[j] = o;
    } else if (i < P.n_so3 + P.n_r3 + P.n_ba + P.n_bg) {
      const bool is_a = i < P.n_so3 + P.n_r3 + P.n_ba;
      const int j = is_a ? i - P.n_so3 - P.n_r3 : i - P.n_so3 - P.n_r3 - P.n_ba;
      const double4 k = is_a ? cur.ba[j] : cur.bg[j];
      const int c = is_a ? P.ba_col[j] : P.bg_col[j];
      const double lim = is_a ? max_ba : max_bg;
      double4 o = k;
      if (c >= 0) {
        // box constraints of SetFixedParams (impl.h:206-251), enforced by projection
        o = make_double4(fmin(fmax(k.x + delta[c], -lim), lim), fmin(fmax(k.y + delta[c + 1], -lim), lim), fmin(fmax(k.z + delta[c + 2], -lim), lim), 0.0);
        step += (o.x - k.x) * (o.x - k.x) + (o.y - k.y) * (o.y - k.y) + (o.z - k.z) * (o.z - k.z);
        xsq += k.x * k.x + k.y * k.y + k.z * k.z;
      }
      if (is_a) cand.ba[j] = o; else cand.bg[j] = o;
    } else {
      double gl[G_COUNT];
      for (int k = 0; k < G_COUNT; ++k) gl[k] = cur.glob[k];
      if (P.col_tic >= 0) {
        Q4 dq; V3 dt;
        se3_exp_dev(delta + P.col_tic, dq, dt);
        const Q4 q = q4(gl[0], gl[1], gl[2], gl[3]);
        const Q4 r = qnormalized(qmul(q, dq));
        const V3 t = v3(gl[4], gl[5], gl[6]) + qrot(q, dt);
        const double nv[7] = {r.x, r.y, r.z, r.w, t.x, t.y, t.z};
        for (int k = 0; k < 7; ++k) { step += (nv[k] - gl[k]) * (nv[k] - gl[k]); xsq += gl[k] * gl[k]; gl[k] = nv[k]; }
      }
      if (P.col_g >= 0) for (int k = 0; k < 3; ++k) { const double d = delta[P.col_g + k]; step += d * d; xsq += gl[G_GRAV + k] * gl[G_GRAV + k]; gl[G_GRAV + k] += d; }
      if (P.col_ld >= 0) { const double d = delta[P.col_ld]; step += d * d; xsq += gl[G_LD] * gl[G_LD]; gl[G_LD] += d; }
      for (int k = 0; k < G_COUNT; ++k) cand.glob[k] = gl[k];
    }
  }
  for (int o = 16; o > 0; o >>= 1) { step += __shfl_xor_sync(0xffffffffu, step, o); xsq += __shfl_xor_sync(0xffffffffu, xsq, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(scal + SC_STEP_SQ, step); atomicAdd(scal + SC_X_SQ, xsq); }
}

int pick_panel(const DeviceProblem& P, size_t& smem_bytes) {
  const int nbp = P.nb + 1, CL = P.ldb + nbp;
  const size_t fixed = (size_t)(nbp * nbp + nbp + 2) * sizeof(double) + (size_t)(P.kd * (P.kd + 1) / 2 + 8) * sizeof(unsigned short);
  int PB = 256;
  for (;;) {
    const size_t win = (size_t)(PB + P.kd) * CL * sizeof(double);
    const size_t back = (size_t)(PB + P.kd) * (P.ldb + 1) * sizeof(double);
    smem_bytes = (win > back ? win : back) + fixed + 64;
    if (smem_bytes <= 200 * 1024 || PB <= 8) break;
    PB /= 2;
  }
  return PB;
}

}  // namespace

size_t solve_workspace_doubles(const DeviceProblem& P) {
  const int n = P.nk + P.nb;
  return (size_t)P.nk * P.ldb + (size_t)P.nk * (P.nb + 1) + 2 * (size_t)n + (size_t)P.nk + 64;
}

void launch_compute_scale(const DeviceProblem& P, double* scale, int jacobi, double* scal, cudaStream_t st) {
  const int n = P.nk + P.nb;
  int grid = (n + 255) / 256; if (grid > 148) grid = 148; if (grid < 1) grid = 1;
  scale_kernel<<<grid, 256, 0, st>>>(P, scale, jacobi, scal);
  count_launch();
}

void launch_solve(const DeviceProblem& P, const double* scale, SolveParams sp, double* workspace, double* delta, double* scal, cudaStream_t st) {
  size_t smem = 0;
  const int PB = pick_panel(P, smem);
  static size_t configured = 0;
  if (smem > configured) { cudaFuncSetAttribute(solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = smem; }
  solve_kernel<<<1, SOLVE_THREADS, smem, st>>>(P, scale, sp, workspace, delta, scal, PB);
  count_launch();
}

void launch_update(const DeviceProblem& P, const DeviceState& cur, const DeviceState& cand, const double* delta, double max_ba, double max_bg, double* scal, cudaStream_t st) {
  const int total = P.n_so3 + P.n_r3 + P.n_ba + P.n_bg + 1;
  update_kernel<<<(total + 127) / 128, 128, 0, st>>>(P, cur, cand, delta, max_ba, max_bg, scal);
  count_launch();
}

}  // namespace icc
