// Accelerometer + gyroscope residuals, analytic Jacobians and J^T J / J^T r reduction: the IMU items of the persistent, TMEM-parked evaluation kernel (icc_eval_tmem.cu, sm_100a).
//
// Same arithmetic contract as imu_kernel<true, 0> of icc_eval.cu (reference: AccelerationCostFunctorSplit<6> / GyroCostFunctorSplit<6>,
// basalt_spline/ceres_calib_split_residuals.h:52-93,133-169, under Ceres autodiff + LieLocalParameterization) for the column set of
// the hot CLI's stage 1 with fixed biases; the wider column sets (bias knots, IMU intrinsics, time offset) stay on imu_kernel.
// Machine mapping = icc_vision_tmem.cu:
//  * one persistent CTA per SM, 12 warps at <= 168 registers; the 15 + 6 FP64 tensor-core accumulator fragments of the two tiles
//    (accelerometer 40 columns, gyroscope 24) are parked in tensor memory (tcgen05.st / tcgen05.ld) during the SIMT passes;
//  * the three rows of a sensor are evaluated in ONE pass (icc_imu_rows.cuh): row 0 goes to the warp's shared-memory tile, rows 1
//    and 2 are parked in TMEM (43 / 38 doubles) and expanded into the same tile after the previous row has been contracted;
//  * the samples of all knot-interval cells form ONE packed stream (cells padded to a multiple of 4 = one m8n8k4 k-step) cut into
//    equal runs of 32-lane chunks, one run per warp: a 50-sample cell (1 kHz, dt = 0.05 s) no longer costs two passes with 18 of 32
//    lanes idle in the second; a chunk may straddle two cells (two staged windows, lanes pick theirs).
#pragma once
#include "icc_imu_rows.cuh"
#include "icc_kernels.h"
#include "icc_tile_common.cuh"
#include "icc_tmem_gen.cuh"

namespace icc {
namespace tmi {

constexpr int IW = 12;                       // warps per CTA
constexpr int LDT = TILE_LD, TCOLS = TILE_COLS;
constexpr int TM_ACCA = 0, TM_ACCG = 60, TM_PARK = 84;

struct ImuSlot {
  ImuWin win[2];
  int gidxA[2][TCOLS];                       // accelerometer tile column -> solver column; -1 constant / padding, -2 residual
  int gidxG[2][24];
  int cinfo[2][8];                           // per staged cell: padded stream offset, first sample, sample count, s_so3, s_r3, s_ba, s_bg
};
struct ImuStageArgs { const double4* so3; const double4* r3; const double4* ba; const double4* bg; const int* so3_col; const int* r3_col; int col_g; };

__device__ __noinline__ void stage_cell(ImuSlot* slot, int w, ImuCellP C, ImuStageArgs A) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  ImuWin& W = slot->win[w];
  double4 k = make_double4(0, 0, 0, 1);
  if (lane < 6) k = A.so3[C.s_so3 + lane];
  const Q4 qa = q4(k.x, k.y, k.z, k.w);
  const Q4 qb = q4(__shfl_down_sync(0xffffffffu, k.x, 1), __shfl_down_sync(0xffffffffu, k.y, 1), __shfl_down_sync(0xffffffffu, k.z, 1), __shfl_down_sync(0xffffffffu, k.w, 1));
  if (lane < 5) stage_frame_increment(W.f, lane, qa, qb);
  if (lane == 0) {
    W.f.q0 = qa; W.f.u_so3 = 0.0; W.f.u_r3 = 0.0;
    slot->cinfo[w][0] = C.poff; slot->cinfo[w][1] = C.i0; slot->cinfo[w][2] = C.n; slot->cinfo[w][3] = C.s_so3; slot->cinfo[w][4] = C.s_r3; slot->cinfo[w][5] = C.s_ba; slot->cinfo[w][6] = C.s_bg;
  }
  if (lane >= 8 && lane < 14) { const double4 p = A.r3[C.s_r3 + lane - 8]; W.f.p[lane - 8] = v3(p.x, p.y, p.z); }
  if (lane >= 16 && lane < 19) { const double4 b = A.ba[C.s_ba + lane - 16]; W.ba[lane - 16] = v3(b.x, b.y, b.z); }
  if (lane >= 20 && lane < 23) { const double4 b = A.bg[C.s_bg + lane - 20]; W.bg[lane - 20] = v3(b.x, b.y, b.z); }
  for (int c = lane; c < TCOLS; c += 32) {
    int g = -1;
    if (c < 18) { const int b = A.so3_col[C.s_so3 + c / 3]; g = b < 0 ? -1 : b + c % 3; }
    else if (c < 36) { const int b = A.r3_col[C.s_r3 + (c - 18) / 3]; g = b < 0 ? -1 : b + (c - 18) % 3; }
    else if (c < 39) g = A.col_g < 0 ? -1 : A.col_g + (c - 36);
    else if (c == ACC_RES_COL) g = -2;
    slot->gidxA[w][c] = g;
    if (c < 24) slot->gidxG[w][c] = c < 18 ? g : (c == GYR_RES_COL ? -2 : -1);
  }
  __syncwarp();
}

// accelerometer tile part (rows [4 k0, 4 k1) of the chunk belong to one cell): rows 0, 1, 2 contracted one after the other
ICC_D void accel_part(const NeLayout& L, const ImuConst& K, double* __restrict__ tile, int k0, int k1, bool first, bool last, const int* __restrict__ gidx, uint32_t ta, int lane) {
  double acc[30];
  if (first) {
#pragma unroll
    for (int i = 0; i < 30; ++i) acc[i] = 0.0;
  } else {
    tmem_ld_d30(ta + TM_ACCA, acc);
  }
  tile_syrk<5>(tile, k0, k1, acc);
  const bool mine = lane >= 4 * k0 && lane < 4 * k1;
#pragma unroll 1
  for (int k = 1; k <= 2; ++k) {
    __syncwarp();
    {
      double so3[18], tail[7];                                   // tail = [q (4) | r1 | r2 | u_r3]
      tmem_ld_d18(ta + TM_PARK + (k == 1 ? 0 : 36), so3);
      tmem_ld_d7(ta + TM_PARK + 72, tail);
      if (mine) imu_accel_row_expand(so3, q4(tail[0], tail[1], tail[2], tail[3]), k == 1 ? tail[4] : tail[5], tail[6], K, k, tile + lane, LDT);
    }
    __syncwarp();
    tile_syrk<5>(tile, k0, k1, acc);
  }
  __syncwarp();
  if (last) tile_flush<5, 4>(L, gidx, acc);
  else tmem_st_d30(ta + TM_ACCA, acc);
}

ICC_D void gyro_part(const NeLayout& L, double* __restrict__ tile, int k0, int k1, bool first, bool last, const int* __restrict__ gidx, uint32_t ta, int lane) {
  double acc[12];
  if (first) {
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = 0.0;
  } else {
    tmem_ld_d12(ta + TM_ACCG, acc);
  }
  tile_syrk<3>(tile, k0, k1, acc);
  const bool mine = lane >= 4 * k0 && lane < 4 * k1;
#pragma unroll 1
  for (int k = 1; k <= 2; ++k) {
    __syncwarp();
    {
      double so3[18], tail[2];
      tmem_ld_d18(ta + TM_PARK + (k == 1 ? 0 : 36), so3);
      tmem_ld_d2(ta + TM_PARK + 72, tail);
      if (mine) {
        double* row = tile + lane;
#pragma unroll
        for (int c = 0; c < 18; ++c) row[c * LDT] = so3[c];
        row[GYR_RES_COL * LDT] = k == 1 ? tail[0] : tail[1];
#pragma unroll
        for (int c = GYR_RES_COL + 1; c < 24; ++c) row[c * LDT] = 0.0;
      }
    }
    __syncwarp();
    tile_syrk<3>(tile, k0, k1, acc);
  }
  __syncwarp();
  if (last) tile_flush<3, 2>(L, gidx, acc);
  else tmem_st_d12(ta + TM_ACCG, acc);
}

ICC_D void init_const(ImuConst* K, const DeviceProblem& P, const DeviceState& S) {
  if (threadIdx.x == 64) {
    const double* ai = S.glob + G_ACC_INTR; const double* gi = S.glob + G_GYR_INTR;
    // misalignment * scale matrices (utils/types.h:226-246)
    const double Ma[9] = {ai[3], -ai[0] * ai[4], ai[1] * ai[5], 0.0, ai[4], -ai[2] * ai[5], 0.0, 0.0, ai[5]};
    const double Mg[9] = {gi[6], -gi[0] * gi[7], gi[1] * gi[8], gi[3] * gi[6], gi[7], -gi[2] * gi[8], -gi[4] * gi[6], gi[5] * gi[7], gi[8]};
    for (int i = 0; i < 9; ++i) { K->Ma[i] = Ma[i]; K->Mg[i] = Mg[i]; }
    K->grav = v3(S.glob[G_GRAV], S.glob[G_GRAV + 1], S.glob[G_GRAV + 2]);
    K->w_acc = P.w_acc; K->w_gyr = P.w_gyr; K->idt2 = P.inv_r3_dt * P.inv_r3_dt; K->inv_so3_dt = P.inv_so3_dt;
  }
}

// One item = one contiguous run of 32-lane chunks of the packed IMU sample stream, processed by one warp.
ICC_D void run_item(const DeviceProblem& P, const DeviceState& S, const ImuConst* K, ImuSlot* slot, double* __restrict__ tile, uint32_t ta, const NeLayout& L, const VisItem it, double* __restrict__ res_out, int lane, int rounds, int bar_id, int bar_n) {
  const bool lock = rounds > 0;
  ImuStageArgs A; A.so3 = S.so3; A.r3 = S.r3; A.ba = S.ba; A.bg = S.bg; A.so3_col = P.so3_col; A.r3_col = P.r3_col; A.col_g = P.col_g;
  const double dto = S.glob[G_TOFF];          // time-offset increment [s] (0 unless the extension has been optimised)
  const double ba_rate = 1e9 / double(P.dt_ba_ns), bg_rate = 1e9 / double(P.dt_bg_ns);
  {
    int c = it.vf0, pos = it.pos_begin, cur = 0;
    stage_cell(slot, cur, P.icells[c], A);
    int endC = min(P.icells[c + 1].poff, it.pos_end);
    bool fresh = true;
    while (pos < it.pos_end) {
      if (rounds > 0) { asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(bar_n) : "memory"); --rounds; }   // loose lockstep of the warps of a group (instruction-cache locality)
      if (pos == endC) {
        ++c;
        stage_cell(slot, cur, P.icells[c], A);
        endC = min(P.icells[c + 1].poff, it.pos_end);
        fresh = true;
      }
      const int nA = min(32, endC - pos);
      int nB = 0, endB = 0;
      if (nA < 32 && endC < it.pos_end) {
        endB = min(P.icells[c + 2].poff, it.pos_end);
        nB = min(32 - nA, endB - endC);
        stage_cell(slot, cur ^ 1, P.icells[c + 1], A);
      }
      const int n = nA + nB;
      const bool inA = lane < nA;
      const int w = inA ? cur : cur ^ 1;
      const int rel = inA ? pos + lane - slot->cinfo[w][0] : lane - nA;
      const bool act = lane < n && rel < slot->cinfo[w][2];
      const int i = slot->cinfo[w][1] + rel;
      const bool lastA = pos + nA == endC, lastB = nB ? endC + nB == endB : false;
      // ---- accelerometer: SIMT pass, then three tensor-core passes per cell part ----------------------------------------
      {
        double park[ACC_PARK];
        double* row0 = tile + lane;
        if (act) {
          const int64_t st = P.imu_t_ns[i];
          // CalcTimes (impl.h:763-788): u = (st % dt) / dt with the segment index known from the cell
          const double u_so3 = double(st - (int64_t)slot->cinfo[w][3] * P.dt_so3_ns) / double(P.dt_so3_ns) + dto * P.inv_so3_dt;
          const double u_r3 = double(st - (int64_t)slot->cinfo[w][4] * P.dt_r3_ns) / double(P.dt_r3_ns) + dto * P.inv_r3_dt;
          const double u_ba = double(st - (int64_t)slot->cinfo[w][5] * P.dt_ba_ns) / double(P.dt_ba_ns) + dto * ba_rate;
          double ra[3];
          imu_accel_rows(slot->win[w], *K, u_so3, u_r3, u_ba, v3(P.imu_acc[3 * i], P.imu_acc[3 * i + 1], P.imu_acc[3 * i + 2]), row0, LDT, park, ra);
          if (res_out) { res_out[P.n_res_vis + 3 * i] = ra[0]; res_out[P.n_res_vis + 3 * i + 1] = ra[1]; res_out[P.n_res_vis + 3 * i + 2] = ra[2]; }
        } else {
#pragma unroll
          for (int cc = 0; cc <= ACC_RES_COL; ++cc) row0[cc * LDT] = 0.0;
#pragma unroll
          for (int cc = 0; cc < ACC_PARK; ++cc) park[cc] = 0.0;
        }
        tmem_st_d43(ta + TM_PARK, park);
      }
      tmem_wait_st();
      __syncwarp();
#pragma unroll 1
      for (int part = 0; part < (nB ? 2 : 1); ++part)
        accel_part(L, *K, tile, part ? nA >> 2 : 0, part ? n >> 2 : nA >> 2, part ? true : fresh, part ? lastB : lastA, slot->gidxA[part ? cur ^ 1 : cur], ta, lane);
      tmem_wait_st();
      __syncwarp();
      // ---- gyroscope ------------------------------------------------------------------------------------------------------
      if (lock) asm volatile("bar.sync %0, %1;" :: "r"(bar_id + 6), "r"(bar_n) : "memory");
      {
        double park[GYR_PARK];
        double* row0 = tile + lane;
        if (act) {
          const int64_t st = P.imu_t_ns[i];
          const double u_so3 = double(st - (int64_t)slot->cinfo[w][3] * P.dt_so3_ns) / double(P.dt_so3_ns) + dto * P.inv_so3_dt;
          const double u_bg = double(st - (int64_t)slot->cinfo[w][6] * P.dt_bg_ns) / double(P.dt_bg_ns) + dto * bg_rate;
          double rg[3];
          imu_gyro_rows(slot->win[w], *K, u_so3, u_bg, v3(P.imu_gyr[3 * i], P.imu_gyr[3 * i + 1], P.imu_gyr[3 * i + 2]), row0, LDT, park, rg);
          if (res_out) { res_out[P.n_res_vis + P.n_res_acc + 3 * i] = rg[0]; res_out[P.n_res_vis + P.n_res_acc + 3 * i + 1] = rg[1]; res_out[P.n_res_vis + P.n_res_acc + 3 * i + 2] = rg[2]; }
        } else {
#pragma unroll
          for (int cc = 0; cc < 24; ++cc) row0[cc * LDT] = 0.0;
#pragma unroll
          for (int cc = 0; cc < GYR_PARK; ++cc) park[cc] = 0.0;
        }
        tmem_st_d38(ta + TM_PARK, park);
      }
      tmem_wait_st();
      __syncwarp();
#pragma unroll 1
      for (int part = 0; part < (nB ? 2 : 1); ++part)
        gyro_part(L, tile, part ? nA >> 2 : 0, part ? n >> 2 : nA >> 2, part ? true : fresh, part ? lastB : lastA, slot->gidxG[part ? cur ^ 1 : cur], ta, lane);
      fresh = false;
      if (nB) { ++c; cur ^= 1; endC = endB; }
      tmem_wait_st();
      __syncwarp();
      pos += n;
    }
  }
  while (rounds-- > 0) { asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(bar_n) : "memory"); asm volatile("bar.sync %0, %1;" :: "r"(bar_id + 6), "r"(bar_n) : "memory"); }
}

}  // namespace tmi
}  // namespace icc
