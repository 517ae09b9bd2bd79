// Persistent TMEM-parked Jacobian evaluation kernel (sm_100a): every warp runs one contiguous run of the packed corner stream
// (vision items, icc_vision_tmem.cuh) or of the packed IMU sample stream (IMU items, icc_imu_tmem.cuh).
//
// One kernel body, TWO launches per evaluation (vision items, then IMU items).  Running both item types in a single launch was
// measured and rejected: with twelve warps per SM spread over the code of both phases the instruction-cache hit rate fell to 64 %
// (`no_instruction` became the top stall, config 4: 300-380 us against 154 + 88 us for the two launches) -- the unrolled FP64
// recursions make this kernel ~150 KB of SASS, and a launch that only walks one phase keeps its working set to that phase.
// Small problems (fewer than a few chunks per warp) stay on the one-warp-per-frame / per-cell kernels of icc_eval.cu, which run
// side by side on two streams and have the shorter single-item latency (launch_eval decides).
//   reference: RSReprojectionCostFunctorSplit<6> / AccelerationCostFunctorSplit<6> / GyroCostFunctorSplit<6>
//   (basalt_spline/ceres_calib_split_residuals.h:319-402, 52-93, 133-169) under Ceres autodiff + LieLocalParameterization.
#include "icc_kernels.h"
#include "icc_vision_tmem.cuh"
#include <cstdlib>
#include "icc_imu_tmem.cuh"

namespace icc {

void count_launch();

namespace {

constexpr int EW = 12;                                   // warps per CTA: 3 per scheduler at <= 168 registers
constexpr int ETM_PER_WARP = 170;                        // 512 TMEM columns / 3 warps per lane quarter
static_assert(tmv::VW == EW && tmi::IW == EW, "one CTA shape");
constexpr size_t SLOT_BYTES = sizeof(tmi::ImuSlot) > sizeof(tmv::WarpSlot) ? sizeof(tmi::ImuSlot) : sizeof(tmv::WarpSlot);
constexpr size_t CONST_BYTES = ((sizeof(VisConst) + sizeof(ImuConst) + 15) / 16) * 16;

template <int MODEL>
__global__ void __launch_bounds__(EW * 32, 1) eval_tmem_kernel(DeviceProblem P, DeviceState S, double* __restrict__ res_out, int rounds, int groups, int stagger_ns) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  VisConst* KV = reinterpret_cast<VisConst*>(smem_raw);
  ImuConst* KI = reinterpret_cast<ImuConst*>(smem_raw + sizeof(VisConst));
  uint32_t* tm_base_s = reinterpret_cast<uint32_t*>(smem_raw + CONST_BYTES);
  unsigned char* slots = smem_raw + CONST_BYTES + 16;
  double* tiles = reinterpret_cast<double*>(slots + EW * SLOT_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* tile = tiles + warp * (TILE_COLS * TILE_LD);

  if (warp == 0) tmem_alloc((uint32_t)__cvta_generic_to_shared(tm_base_s), 512);
  tmv::init_const(KV, P, S, MODEL);
  tmi::init_const(KI, P, S);
  for (int i = lane; i < TILE_COLS * TILE_LD; i += 32) tile[i] = 0.0;
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  // lane quarter of this warp (hardware rule: warp w may touch TMEM lanes 32 (w % 4) ..+31), column group by warp / 4
  const uint32_t ta = *tm_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * ETM_PER_WARP);
  NeLayout L; L.ne = P.ne; L.off_E = P.ne_off_E; L.off_C = P.ne_off_C; L.off_g = P.ne_off_g; L.off_cost = P.ne_off_cost; L.nk = P.nk; L.nb = P.nb; L.ldb = P.ldb;

  const int gw = warp * gridDim.x + blockIdx.x;           // global warp: consecutive runs land on different SMs
  // lockstep groups: the whole CTA (groups == 1), or the three sets of four warps that sit on the four sub-cores together (warp / 4),
  // each with its own named barriers and started `stagger_ns` apart, so that one group's tensor-core phase meets the others' SIMT phases
  const int grp = groups == 3 ? warp >> 2 : groups == 4 ? (warp & 3) : groups == 6 ? warp % 6 : groups == 2 ? (warp & 1) : 0, bar_id = 1 + grp, bar_n = EW * 32 / (groups > 0 ? groups : 1);
  if (groups > 1 && stagger_ns > 0) for (int g = 0; g < grp; ++g) __nanosleep(stagger_ns);
  if (P.n_vitems > 0) {
    if (gw < P.n_vitems) tmv::run_item<MODEL>(P, S, KV, reinterpret_cast<tmv::WarpSlot*>(slots + warp * SLOT_BYTES), tile, ta, L, P.vitems[gw], res_out, lane, rounds, bar_id, bar_n);
    else for (int r = 0; r < rounds; ++r) asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(bar_n) : "memory");
  } else if (P.n_iitems > 0) {           // (a launch carries one item type: launch_eval_tmem)
    if (gw < P.n_iitems) tmi::run_item(P, S, KI, reinterpret_cast<tmi::ImuSlot*>(slots + warp * SLOT_BYTES), tile, ta, L, P.iitems[gw], res_out, lane, rounds, bar_id, bar_n);
    else for (int r = 0; r < rounds; ++r) { asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(bar_n) : "memory"); asm volatile("bar.sync %0, %1;" :: "r"(bar_id + 6), "r"(bar_n) : "memory"); }
  }
  tmem_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*tm_base_s, 512);
}

template <int MODEL>
int launch_model(const DeviceProblem& P, const DeviceState& S, double* residuals_out, int grid, size_t smem, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(eval_tmem_kernel<MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
    attr_done = true;
  }
  // loose lockstep: warps start every chunk together (one named barrier per chunk), so that they walk the same ~80 KB of unrolled code at
  // the same time (instruction-cache hit rate 81 % -> better).  Measured on config 4 (tests/stagger_probe.py), vision / IMU launch in us:
  // free running 153 / 88, whole CTA in step 149.5 / 84.1, three groups of four warps (one warp per sub-core each) 143.7 / 85.0, four
  // groups of three (the warps of one sub-core) 147.7 / 80.6, six pairs 147.5 / 94.7; starting the groups a few us apart only hurts.
  // Hence vision items run as 3 groups, IMU items as 4 (ICC_TMEM_GROUPS overrides, ICC_TMEM_STAGGER_NS delays group g by g x ns).
  const int rounds = getenv("ICC_TMEM_NO_LOCKSTEP") ? 0 : P.n_vitems > 0 ? (P.n_vchunks + P.n_vitems - 1) / P.n_vitems : P.n_iitems > 0 ? (P.n_ichunks + P.n_iitems - 1) / P.n_iitems : 0;
  const char* eg = getenv("ICC_TMEM_GROUPS"); const char* es = getenv("ICC_TMEM_STAGGER_NS");
  const int gq = eg ? atoi(eg) : (P.n_vitems > 0 ? 3 : 4), groups = (gq == 2 || gq == 3 || gq == 4 || gq == 6) ? gq : 1, stagger_ns = es ? atoi(es) : 0;
  eval_tmem_kernel<MODEL><<<grid, EW * 32, smem, st>>>(P, S, residuals_out, rounds, groups, stagger_ns);
  return 0;
}

}  // namespace

int eval_tmem_warps() { return EW; }

// P.vitems[k] / P.iitems[k] is the run of global warp k (icc_api.cu cuts each stream into at most sm_count * eval_tmem_warps() runs);
// vision and IMU items go out as separate launches of the same kernel.
int launch_eval_tmem(const DeviceProblem& P, const DeviceState& S, double* residuals_out, int sm_count, cudaStream_t st) {
  const size_t smem = CONST_BYTES + 16 + EW * SLOT_BYTES + (size_t)EW * TILE_COLS * TILE_LD * sizeof(double);
  for (int pass = 0; pass < 2; ++pass) {
    DeviceProblem Q = P;
    if (pass == 0) Q.n_iitems = 0; else Q.n_vitems = 0;
    const int n = pass == 0 ? Q.n_vitems : Q.n_iitems;
    if (n <= 0) continue;
    const int grid = (n + EW - 1) / EW < sm_count ? (n + EW - 1) / EW : sm_count;
    int e = 1;
    switch (P.model) {   // one instantiation per camera model: only that model's projection code is resident in the instruction cache
      case CAM_PINHOLE: e = launch_model<CAM_PINHOLE>(Q, S, residuals_out, grid, smem, st); break;
      case CAM_PINHOLE_RADTAN: e = launch_model<CAM_PINHOLE_RADTAN>(Q, S, residuals_out, grid, smem, st); break;
      case CAM_FISHEYE: e = launch_model<CAM_FISHEYE>(Q, S, residuals_out, grid, smem, st); break;
      case CAM_FOV: e = launch_model<CAM_FOV>(Q, S, residuals_out, grid, smem, st); break;
      case CAM_DIVISION_UNDISTORTION: e = launch_model<CAM_DIVISION_UNDISTORTION>(Q, S, residuals_out, grid, smem, st); break;
      case CAM_DOUBLE_SPHERE: e = launch_model<CAM_DOUBLE_SPHERE>(Q, S, residuals_out, grid, smem, st); break;
      case CAM_EXTENDED_UNIFIED: e = launch_model<CAM_EXTENDED_UNIFIED>(Q, S, residuals_out, grid, smem, st); break;
      default: return 1;
    }
    if (e) return 1;
    count_launch();
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace icc
