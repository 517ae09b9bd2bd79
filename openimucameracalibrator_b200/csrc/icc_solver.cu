// Damped normal-equation solve + manifold update kernels (sm_100a).
//
// Replaces what ceres::Solve does between two Jacobian evaluations for the reference's spline problem
// (SplineTrajectoryEstimator::Optimize, include/OpenCameraCalibrator/core/spline_trajectory_estimator.impl.h:254-276):
//   * Jacobi column scaling + Levenberg-Marquardt diagonal  (Ceres TrustRegionMinimizer / LevenbergMarquardtStrategy)
//   * SPARSE_NORMAL_CHOLESKY on (J^T J + D^T D)              -> exact LDL^T of the banded (spline knots, time ordered)
//     + bordered (T_i_c, gravity, line delay, bias knots) system: the spline control points are Schur-eliminated first
//   * LieLocalParameterization::Plus on every SO(3) knot and on T_i_c (basalt_spline/ceres_local_param.h:84-92)
//
// The band factorisation is a recurrence along time, so it is cut into P time chunks (substructuring / one-level nested
// dissection): kernel A eliminates every chunk's interior knots in parallel (one CTA per chunk, active band window in
// shared memory, one __syncthreads per eliminated column, LDL^T so no square roots), carrying the couplings to the
// chunk's left separator, the border and the right-hand side; kernel B factors the small reduced system
// {separators (block tridiagonal) + border} in one CTA and back-substitutes it; kernel C back-substitutes the interiors
// in parallel; kernel D forms the step and the model cost change.
#include "icc_device_math.cuh"
#include "icc_kernels.h"

#include <cmath>

namespace icc {

void count_launch();

namespace {

constexpr int NT = 512;          // threads of the factorisation kernels
constexpr int MAX_CHUNKS = 32;

struct SolvePlan {
  int P, w;                       // chunks, separator width (0 when P == 1)
  int a[MAX_CHUNKS], b[MAX_CHUNKS];   // interior column ranges [a, b)
  int nkr, kdr, ldbr;             // reduced (separator) system: columns, half bandwidth, column length
  int nbl;                        // local border rows of a chunk: w + nb + 1 (left separator | border | rhs)
  int WS_A, WS_B;                 // window slots (power of two)
};

// Workspace (doubles): Lb[nk*ldb] | El[nk*nbl] | y[n] | t[nk] | R{bandr[nkr*ldbr] | Er[nkr*nbp] | Cr[nbp*nbp]} | Lbr[nkr*ldbr] | Elr[nkr*nbp] | tr[nkr]
struct SolveWs { double *Lb, *El, *y, *t, *bandr, *Er, *Cr, *Lbr, *Elr, *tr; size_t reduced_doubles, total; };
__host__ __device__ inline SolveWs carve(double* ws, int nk, int nb, int ldb, const SolvePlan& pl) {
  SolveWs w; const int nbp = nb + 1; const size_t n = (size_t)nk + nb;
  w.Lb = ws; w.El = w.Lb + (size_t)nk * ldb; w.y = w.El + (size_t)nk * pl.nbl; w.t = w.y + n;
  w.bandr = w.t + nk; w.Er = w.bandr + (size_t)pl.nkr * pl.ldbr; w.Cr = w.Er + (size_t)pl.nkr * nbp;
  w.reduced_doubles = (size_t)pl.nkr * pl.ldbr + (size_t)pl.nkr * nbp + (size_t)nbp * nbp;
  w.Lbr = w.Cr + (size_t)nbp * nbp; w.Elr = w.Lbr + (size_t)pl.nkr * pl.ldbr; w.tr = w.Elr + (size_t)pl.nkr * nbp;
  w.total = (size_t)(w.tr + pl.nkr - ws) + 16;
  return w;
}

__device__ __forceinline__ int tri_row(int idx) {   // idx = r(r+1)/2 + c, 0 <= c <= r  ->  r
  int r = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
  while ((r + 1) * (r + 2) / 2 <= idx) ++r;
  while (r * (r + 1) / 2 > idx) --r;
  return r;
}

// LM diagonal: clamp(diag(S H S), min, max) / radius   (LevenbergMarquardtStrategy::ComputeStep)
__device__ __forceinline__ double lm_d2(const DeviceProblem& P, const double* scale, const SolveParams& sp, int i) {
  const double d = i < P.nk ? P.ne[(int64_t)i * P.ldb] : P.ne[P.ne_off_C + (int64_t)(i - P.nk) * P.nb + (i - P.nk)];
  const double s = scale[i];
  return fmin(fmax(d * s * s, sp.min_diag), sp.max_diag) / sp.radius;
}

// Update descriptors of one column elimination: value = col[srcA] * col[srcB] / pivot is subtracted from
//   window column (j + dc) at offset dst            (dc != 0xFFFF)   — band x band and border x band entries
//   the dense local-border block at offset dst       (dc == 0xFFFF)   — border x border entries
__device__ void build_descriptors(ushort4* desc, int kd, int ldb, int nbl) {
  const int np = kd * (kd + 1) / 2, nbb = nbl * kd, ncc = nbl * (nbl + 1) / 2;
  for (int idx = threadIdx.x; idx < np + nbb + ncc; idx += blockDim.x) {
    ushort4 d;
    if (idx < np) { const int r0 = tri_row(idx), c0 = idx - r0 * (r0 + 1) / 2; const int r = r0 + 1, c = c0 + 1; d = make_ushort4(r, c, c, r - c); }
    else if (idx < np + nbb) { const int k = idx - np, b = k / kd, c = k % kd + 1; d = make_ushort4(ldb + b, c, c, ldb + b); }
    else { const int k = idx - np - nbb, b = tri_row(k), c = k - b * (b + 1) / 2; d = make_ushort4(ldb + b, ldb + c, 0xFFFF, b * nbl + c); }
    desc[idx] = d;
  }
}

// LDL^T elimination of columns [j_begin, j_end) of a banded matrix with dense "local border" rows, through a circular
// shared-memory window of WS column slots (each CL = ldb + nbl doubles).  load(col, e) returns the (scaled, damped)
// original entry e of column col (0 outside the matrix); store(col, e, v) receives every finished column (unscaled
// LDL^T storage: entry 0 = pivot D_j, others = L_ij D_j).  On return the window still holds the updated columns
// [j_end, j_end + kd].  Returns false on a non-positive pivot.
template <class Load, class Store>
__device__ bool factor_range(double* W, double* Cl, const ushort4* desc, int T, int j_begin, int j_end, int kd, int CL, int WS, Load load, Store store) {
  const int tid = threadIdx.x, nt = blockDim.x, PB = WS - kd - 1, mask = WS - 1;
  for (int j0 = j_begin; j0 < j_end; j0 += PB) {
    const int first = j0 == j_begin ? j_begin : j0 + kd + 1, last = j0 + PB + kd + 1;
    for (int idx = tid; idx < (last - first) * CL; idx += nt) { const int col = first + idx / CL, e = idx % CL; W[(size_t)(col & mask) * CL + e] = load(col, e); }
    __syncthreads();
    const int jend = min(j0 + PB, j_end);
    for (int j = j0; j < jend; ++j) {
      const double* cj = W + (size_t)(j & mask) * CL;
      const double piv = cj[0];
      if (!(piv > 0.0) || !isfinite(piv)) return false;       // uniform: every thread reads the same pivot
      const double inv = 1.0 / piv;
      for (int idx = tid; idx < T; idx += nt) {
        const ushort4 d = desc[idx];
        const double v = cj[d.x] * cj[d.y] * inv;
        double* dst = d.z == 0xFFFF ? Cl + d.w : W + (size_t)((j + d.z) & mask) * CL + d.w;
        *dst -= v;
      }
      __syncthreads();
    }
    for (int idx = tid; idx < (jend - j0) * CL; idx += nt) { const int col = j0 + idx / CL, e = idx % CL; store(col, e, W[(size_t)(col & mask) * CL + e]); }
    __syncthreads();
  }
  return true;
}

// ---- kernel A: eliminate the interior knots of every time chunk ---------------------------------------------------
__global__ void __launch_bounds__(NT) chunk_factor_kernel(DeviceProblem P, SolvePlan pl, const double* __restrict__ scale, SolveParams sp, double* wsp, double* scal) {
  extern __shared__ __align__(16) double sm[];
  const int c = blockIdx.x, nk = P.nk, nb = P.nb, kd = P.kd, ldb = P.ldb, w = pl.w, nbl = pl.nbl, nbp = nb + 1;
  const int a = pl.a[c], b = pl.b[c];
  const bool has_left = c > 0, has_right = c < pl.P - 1;
  const int CL = ldb + nbl, WS = pl.WS_A;
  double* W = sm;
  double* Cl = W + (size_t)WS * CL;
  ushort4* desc = reinterpret_cast<ushort4*>(Cl + nbl * nbl);
  const int T = kd * (kd + 1) / 2 + nbl * kd + nbl * (nbl + 1) / 2;
  const double* band = P.ne; const double* E = P.ne + P.ne_off_E; const double* g = P.ne + P.ne_off_g;
  SolveWs ws = carve(wsp, nk, nb, ldb, pl);
  build_descriptors(desc, kd, ldb, nbl);
  for (int i = threadIdx.x; i < nbl * nbl; i += blockDim.x) Cl[i] = 0.0;
  __syncthreads();
  const int right_end = has_right ? b + w : b;
  auto load = [&](int col, int e) -> double {
    if (col >= right_end) return 0.0;
    if (e < ldb) {
      const int i = col + e;
      if (i >= nk || i >= right_end) return 0.0;                // rows beyond the right separator belong to the next chunk
      double v = band[(int64_t)col * ldb + e] * scale[col] * scale[i];
      if (e == 0) v += lm_d2(P, scale, sp, col);
      return v;
    }
    const int lb = e - ldb;
    if (lb < w) {                                               // coupling to the left separator (stored transposed in H)
      if (!has_left || col >= b) return 0.0;
      const int s = a - w + lb, off = col - s;
      return off <= kd ? band[(int64_t)s * ldb + off] * scale[s] * scale[col] : 0.0;
    }
    if (lb < w + nb) return E[(int64_t)col * nb + (lb - w)] * scale[col] * scale[nk + lb - w];
    return -g[col] * scale[col];
  };
  auto store = [&](int col, int e, double v) { if (e < ldb) ws.Lb[(int64_t)col * ldb + e] = v; else ws.El[(int64_t)col * nbl + (e - ldb)] = v; };
  const bool ok = factor_range(W, Cl, desc, T, a, b, kd, CL, WS, load, store);
  if (!ok) { if (threadIdx.x == 0) scal[SC_OK] = -1.0; return; }
  // ---- scatter the Schur complement of this chunk into the reduced system ------------------------------------------
  const int mask = WS - 1;
  if (b == a) {   // empty interior (cannot happen with a valid plan) — still need the separator columns in the window
    for (int idx = threadIdx.x; idx < (right_end - b) * CL; idx += blockDim.x) { const int col = b + idx / CL, e = idx % CL; W[(size_t)(col & mask) * CL + e] = load(col, e); }
    __syncthreads();
  }
  const int sl0 = (c - 1) * w, sr0 = c * w;                     // reduced indices of the left / right separator
  if (has_right) {
    for (int idx = threadIdx.x; idx < w * CL; idx += blockDim.x) {
      const int t = idx / CL, e = idx % CL, col = b + t;
      const double v = W[(size_t)(col & mask) * CL + e];
      if (v == 0.0) continue;
      if (e < ldb) { if (t + e < w) atomicAdd(ws.bandr + (int64_t)(sr0 + t) * pl.ldbr + e, v); }
      else {
        const int lb = e - ldb;
        if (lb < w) { if (has_left) atomicAdd(ws.bandr + (int64_t)(sl0 + lb) * pl.ldbr + (sr0 + t - sl0 - lb), v); }
        else atomicAdd(ws.Er + (int64_t)(sr0 + t) * nbp + (lb - w), v);
      }
    }
  }
  for (int idx = threadIdx.x; idx < nbl * (nbl + 1) / 2; idx += blockDim.x) {
    const int b1 = tri_row(idx), b2 = idx - b1 * (b1 + 1) / 2;   // b1 >= b2, local order [left | border | rhs]
    const double v = Cl[b1 * nbl + b2];
    if (v == 0.0) continue;
    if (b1 < w) { if (has_left) atomicAdd(ws.bandr + (int64_t)(sl0 + b2) * pl.ldbr + (b1 - b2), v); }
    else if (b2 < w) { if (has_left) atomicAdd(ws.Er + (int64_t)(sl0 + b2) * nbp + (b1 - w), v); }
    else atomicAdd(ws.Cr + (int64_t)(b1 - w) * nbp + (b2 - w), v);
  }
}

// ---- kernel B: reduced system {separators + border}: factor, solve ------------------------------------------------
__global__ void __launch_bounds__(NT) reduced_solve_kernel(DeviceProblem P, SolvePlan pl, const double* __restrict__ scale, SolveParams sp, double* wsp, double* scal) {
  extern __shared__ __align__(16) double sm[];
  const int nk = P.nk, nb = P.nb, nbp = nb + 1, nkr = pl.nkr, kdr = pl.kdr, ldbr = pl.ldbr, w = pl.w;
  const int CL = ldbr + nbp, WS = pl.WS_B, tid = threadIdx.x, nt = blockDim.x;
  double* W = sm;
  double* Cs = W + (size_t)WS * CL;                 // nbp x nbp lower, row nb = rhs
  double* xb = Cs + nbp * nbp;                      // nbp
  ushort4* desc = reinterpret_cast<ushort4*>(xb + nbp + 1);
  const int T = kdr * (kdr + 1) / 2 + nbp * kdr + nbp * (nbp + 1) / 2;
  const double* C = P.ne + P.ne_off_C; const double* g = P.ne + P.ne_off_g;
  SolveWs ws = carve(wsp, nk, nb, P.ldb, pl);
  if (scal[SC_OK] < 0.0) { if (tid == 0) { scal[SC_OK] = 0.0; scal[SC_MODEL_CHANGE] = 0.0; } return; }   // a chunk hit a bad pivot
  if (nkr > 0) build_descriptors(desc, kdr, ldbr, nbp);
  for (int idx = tid; idx < nbp * nbp; idx += nt) {
    const int b = idx / nbp, c = idx % nbp;
    double v = 0.0;
    if (c <= b) {
      v = ws.Cr[idx];                                                       // Schur contributions of all chunks
      if (b < nb) { v += C[(int64_t)b * nb + c] * scale[nk + b] * scale[nk + c]; if (b == c) v += lm_d2(P, scale, sp, nk + b); }
      else if (c < nb) v += -g[nk + c] * scale[nk + c];
    }
    Cs[idx] = v;
  }
  __syncthreads();
  bool ok = true;
  if (nkr > 0) {
    auto load = [&](int col, int e) -> double {
      if (col >= nkr) return 0.0;
      if (e < ldbr) return col + e < nkr ? ws.bandr[(int64_t)col * ldbr + e] : 0.0;
      return ws.Er[(int64_t)col * nbp + (e - ldbr)];
    };
    auto store = [&](int col, int e, double v) { if (e < ldbr) ws.Lbr[(int64_t)col * ldbr + e] = v; else ws.Elr[(int64_t)col * nbp + (e - ldbr)] = v; };
    ok = factor_range(W, Cs, desc, T, 0, nkr, kdr, CL, WS, load, store);
  }
  // border: dense LDL^T of the final Schur complement, rhs carried as the last row
  for (int j = 0; j < nb && ok; ++j) {
    const double piv = Cs[j * nbp + j];
    if (!(piv > 0.0) || !isfinite(piv)) { ok = false; break; }
    const double inv = 1.0 / piv;
    const int mrem = nbp - 1 - j;
    for (int idx = tid; idx < mrem * (mrem + 1) / 2; idx += nt) {
      const int r = tri_row(idx), c = idx - r * (r + 1) / 2;
      const int b = j + 1 + r, cc = j + 1 + c;
      Cs[b * nbp + cc] -= Cs[b * nbp + j] * Cs[cc * nbp + j] * inv;
    }
    __syncthreads();
  }
  if (!ok) { if (tid == 0) { scal[SC_OK] = 0.0; scal[SC_MODEL_CHANGE] = 0.0; } return; }
  if (tid < 32) {   // border back-substitution
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
  // separators: t_j = rhs_j - sum_b Elr[j][b] x_b, then descending scatter-form back-substitution (one warp)
  for (int j = tid; j < nkr; j += nt) { const double* le = ws.Elr + (int64_t)j * nbp; double t = le[nb]; for (int b = 0; b < nb; ++b) t -= le[b] * xb[b]; ws.tr[j] = t; }
  __syncthreads();
  if (nkr > 0) {
    const int PB = 64;
    double* Bw = W; double* tw = W + (size_t)(PB + kdr) * ldbr;
    for (int hi = nkr; hi > 0; hi -= PB) {
      const int lo_own = max(0, hi - PB), lo = max(0, lo_own - kdr);
      for (int idx = tid; idx < (hi - lo) * ldbr; idx += nt) Bw[idx] = ws.Lbr[(int64_t)lo * ldbr + idx];
      for (int idx = tid; idx < hi - lo; idx += nt) tw[idx] = ws.tr[lo + idx];
      __syncthreads();
      if (tid < 32) {
        for (int j = hi - 1; j >= lo_own; --j) {
          const int jl = j - lo;
          const double xj = tw[jl] / Bw[(int64_t)jl * ldbr];
          __syncwarp();
          if (tid == 0) tw[jl] = xj;
          for (int r = tid + 1; r <= kdr && r <= jl; r += 32) tw[jl - r] -= Bw[(int64_t)(jl - r) * ldbr + r] * xj;
          __syncwarp();
        }
      }
      __syncthreads();
      for (int idx = tid; idx < hi - lo_own; idx += nt) { const int rj = lo_own + idx; const int k = rj / w, tt = rj % w; ws.y[pl.b[k] + tt] = tw[lo_own - lo + idx]; }
      for (int idx = tid; idx < lo_own - lo; idx += nt) ws.tr[lo + idx] = tw[idx];
      __syncthreads();
    }
  }
  if (tid == 0) scal[SC_OK] = 1.0;
}

// ---- kernel C: back-substitute the chunk interiors in parallel ----------------------------------------------------
__global__ void __launch_bounds__(256) chunk_backsub_kernel(DeviceProblem P, SolvePlan pl, double* wsp, const double* scal) {
  extern __shared__ __align__(16) double sm[];
  if (scal[SC_OK] != 1.0) return;
  const int c = blockIdx.x, nk = P.nk, nb = P.nb, kd = P.kd, ldb = P.ldb, w = pl.w, nbl = pl.nbl, tid = threadIdx.x, nt = blockDim.x;
  const int a = pl.a[c], b = pl.b[c];
  const bool has_left = c > 0, has_right = c < pl.P - 1;
  SolveWs ws = carve(wsp, nk, nb, ldb, pl);
  double* xl = sm;                       // local border solution [left separator | border]
  const int PB = 128;
  double* Bw = xl + nbl;                 // (PB + kd) * ldb
  double* tw = Bw + (size_t)(PB + kd) * ldb;
  for (int i = tid; i < w + nb; i += nt) xl[i] = i < w ? (has_left ? ws.y[a - w + i] : 0.0) : ws.y[nk + i - w];
  __syncthreads();
  for (int j = a + tid; j < b; j += nt) { const double* le = ws.El + (int64_t)j * nbl; double t = le[w + nb]; for (int i = 0; i < w + nb; ++i) t -= le[i] * xl[i]; ws.t[j] = t; }
  __syncthreads();
  const int top = has_right ? b + w : b;   // right separator values are known: they only scatter into the interior
  for (int hi = top; hi > a; hi -= PB) {
    const int lo_own = max(a, hi - PB), lo = max(a, lo_own - kd);
    for (int idx = tid; idx < (hi - lo) * ldb; idx += nt) { const int col = lo + idx / ldb; Bw[idx] = col < b ? ws.Lb[(int64_t)lo * ldb + idx] : 0.0; }
    for (int idx = tid; idx < hi - lo; idx += nt) { const int col = lo + idx; tw[idx] = col < b ? ws.t[col] : ws.y[col]; }
    __syncthreads();
    if (tid < 32) {
      for (int j = hi - 1; j >= lo_own; --j) {
        const int jl = j - lo;
        const double xj = j < b ? tw[jl] / Bw[(int64_t)jl * ldb] : tw[jl];
        __syncwarp();
        if (tid == 0) tw[jl] = xj;
        for (int r = tid + 1; r <= kd && r <= jl; r += 32) { if (j - r < b) tw[jl - r] -= Bw[(int64_t)(jl - r) * ldb + r] * xj; }
        __syncwarp();
      }
    }
    __syncthreads();
    for (int idx = tid; idx < hi - lo_own; idx += nt) { const int col = lo_own + idx; if (col < b) ws.y[col] = tw[lo_own - lo + idx]; }
    for (int idx = tid; idx < lo_own - lo; idx += nt) ws.t[lo + idx] = tw[idx];
    __syncthreads();
  }
}

// ---- kernel D: step in the unscaled space, model cost change = 1/2 (y^T D y + y^T rhs) ----------------------------
__global__ void finish_kernel(DeviceProblem P, SolvePlan pl, const double* __restrict__ scale, SolveParams sp, double* wsp, double* delta, double* scal) {
  if (scal[SC_OK] != 1.0) return;
  const int n = P.nk + P.nb;
  SolveWs ws = carve(wsp, P.nk, P.nb, P.ldb, pl);
  const double* g = P.ne + P.ne_off_g;
  double part = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double yi = ws.y[i], s = scale[i];
    delta[i] = yi * s;
    part += lm_d2(P, scale, sp, i) * yi * yi - yi * g[i] * s;
  }
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(scal + SC_MODEL_CHANGE, 0.5 * part);
}

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
    atomicMax(reinterpret_cast<unsigned long long*>(scal + SC_GRAD_MAX), (unsigned long long)__double_as_longlong(m));   // non-negative doubles order like their bits
  }
}

// SE3::exp (sophus/se3.hpp:761-783), tangent = (upsilon, omega)
__device__ void se3_exp_dev(const double* a, Q4& q, V3& t) {
  const V3 ups = v3(a[0], a[1], a[2]), om = v3(a[3], a[4], a[5]);
  const double th2 = dot(om, om);
  const ExpOut e = so3_exp_jr(om);
  q = e.q;
  if (th2 < kEps * kEps) { t = qrot(q, ups); return; }   // V = so3.matrix() branch
  const V3 c1 = cross(om, ups);                          // V = I + (1-cos t)/t^2 [om]x + (t - sin t)/t^3 [om]x^2
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
      if (c >= 0) {   // box constraints of SetFixedParams (impl.h:206-251), enforced by projection
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

int pow2_at_least(int v) { int p = 1; while (p < v) p <<= 1; return p; }

size_t smem_A(const DeviceProblem& P, const SolvePlan& pl) {
  const int CL = P.ldb + pl.nbl, T = P.kd * (P.kd + 1) / 2 + pl.nbl * P.kd + pl.nbl * (pl.nbl + 1) / 2;
  return ((size_t)pl.WS_A * CL + (size_t)pl.nbl * pl.nbl) * sizeof(double) + (size_t)T * sizeof(ushort4) + 64;
}
size_t smem_B(const DeviceProblem& P, const SolvePlan& pl) {
  const int nbp = P.nb + 1, CL = pl.ldbr + nbp, T = pl.kdr * (pl.kdr + 1) / 2 + nbp * pl.kdr + nbp * (nbp + 1) / 2;
  const size_t win = (size_t)pl.WS_B * CL * sizeof(double), back = (size_t)(64 + pl.kdr) * (pl.ldbr + 1) * sizeof(double);
  return (win > back ? win : back) + ((size_t)nbp * nbp + nbp + 2) * sizeof(double) + (size_t)T * sizeof(ushort4) + 64;
}
size_t smem_C(const DeviceProblem& P, const SolvePlan& pl) { return ((size_t)pl.nbl + (size_t)(128 + P.kd) * (P.ldb + 1)) * sizeof(double) + 64; }

// Chunking of the knot columns.  The elimination cost per interior column is ~ constant, the reduced system has
// (P-1) * kd sequential columns => P ~ sqrt(nk / kd).  Wide borders (bias splines active) keep P = 1: the left-separator
// coupling would make the local border too large for shared memory.
SolvePlan make_plan(const DeviceProblem& P) {
  SolvePlan pl; memset(&pl, 0, sizeof pl);
  const int nk = P.nk, kd = P.kd, nb = P.nb;
  int Pn = 1;
  if (nk > 0 && kd > 0 && nb <= 16) { Pn = (int)std::lround(std::sqrt((double)nk / (double)(kd + 1))); Pn = std::max(1, std::min(Pn, MAX_CHUNKS)); while (Pn > 1 && (nk - (Pn - 1) * kd) / Pn < 2 * (kd + 1)) --Pn; }
  pl.P = Pn; pl.w = Pn > 1 ? kd : 0;
  const int interior_total = nk - (Pn - 1) * pl.w;
  int pos = 0;
  for (int c = 0; c < Pn; ++c) { const int len = interior_total / Pn + (c < interior_total % Pn ? 1 : 0); pl.a[c] = pos; pl.b[c] = pos + len; pos += len + pl.w; }
  pl.nkr = (Pn - 1) * pl.w; pl.kdr = Pn > 1 ? 2 * pl.w - 1 : 0; pl.ldbr = pl.kdr + 1;
  pl.nbl = pl.w + nb + 1;
  pl.WS_A = pow2_at_least(kd + 2); pl.WS_B = pow2_at_least(pl.kdr + 2);
  // grow the windows while they fit comfortably (larger panels amortise the panel load/store)
  while (pl.WS_A < 256) { SolvePlan t = pl; t.WS_A *= 2; if (smem_A(P, t) > 160 * 1024) break; pl = t; }
  while (pl.WS_B < 256) { SolvePlan t = pl; t.WS_B *= 2; if (smem_B(P, t) > 160 * 1024) break; pl = t; }
  return pl;
}

}  // namespace

size_t solve_workspace_doubles(const DeviceProblem& P) {
  const SolvePlan pl = make_plan(P);
  return carve(nullptr, P.nk, P.nb, P.ldb, pl).total;
}

void launch_compute_scale(const DeviceProblem& P, double* scale, int jacobi, double* scal, cudaStream_t st) {
  const int n = P.nk + P.nb;
  int grid = (n + 255) / 256; if (grid > 148) grid = 148; if (grid < 1) grid = 1;
  scale_kernel<<<grid, 256, 0, st>>>(P, scale, jacobi, scal);
  count_launch();
}

void launch_solve(const DeviceProblem& P, const double* scale, SolveParams sp, double* workspace, double* delta, double* scal, cudaStream_t st) {
  const SolvePlan pl = make_plan(P);
  const SolveWs ws = carve(workspace, P.nk, P.nb, P.ldb, pl);
  static size_t cfgA = 0, cfgB = 0, cfgC = 0;
  const size_t sA = smem_A(P, pl), sB = smem_B(P, pl), sC = smem_C(P, pl);
  if (sA > cfgA) { cudaFuncSetAttribute(chunk_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sA); cfgA = sA; }
  if (sB > cfgB) { cudaFuncSetAttribute(reduced_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sB); cfgB = sB; }
  if (sC > cfgC) { cudaFuncSetAttribute(chunk_backsub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sC); cfgC = sC; }
  cudaMemsetAsync(ws.bandr, 0, ws.reduced_doubles * sizeof(double), st);
  if (P.nk > 0) { chunk_factor_kernel<<<pl.P, NT, sA, st>>>(P, pl, scale, sp, workspace, scal); count_launch(); }
  reduced_solve_kernel<<<1, NT, sB, st>>>(P, pl, scale, sp, workspace, scal); count_launch();
  if (P.nk > 0) { chunk_backsub_kernel<<<pl.P, 256, sC, st>>>(P, pl, workspace, scal); count_launch(); }
  const int n = P.nk + P.nb;
  int grid = (n + 255) / 256; if (grid > 148) grid = 148; if (grid < 1) grid = 1;
  finish_kernel<<<grid, 256, 0, st>>>(P, pl, scale, sp, workspace, delta, scal); count_launch();
}

void launch_update(const DeviceProblem& P, const DeviceState& cur, const DeviceState& cand, const double* delta, double max_ba, double max_bg, double* scal, cudaStream_t st) {
  const int total = P.n_so3 + P.n_r3 + P.n_ba + P.n_bg + 1;
  update_kernel<<<(total + 127) / 128, 128, 0, st>>>(P, cur, cand, delta, max_ba, max_bg, scal);
  count_launch();
}

}  // namespace icc
