// kontiki_b200 -- CUDA kernels (sm_100a) and the C ABI of include/kontiki_b200.h.
//
// Kernels per evaluation point:
//   k_pair_prepass   K0: n_knots x 7 (reference layout) -> 64-B knot records; omega_p = log(P_{p-1}^-1 P_p) and its 6x14 ambient Jacobian,
//                    one thread per (pair, direction); clears the status word
//   k_imu<0..3>      gyroscope / accelerometer / position / orientation rows (residual 3 or 1, packed Jacobian 4x3x7); k_short_batch runs all
//                    IMU-like groups of a problem WITHOUT camera rows in one launch
//   k_landmark_ref   reference side of the camera rows, ONCE per landmark reference: X, dX/drho, dX/d(4 knots)
//   k_static_rs      static rolling-shutter camera rows (residual 2, packed Jacobian 2x(28+28+1)); the warp gathers its 32
//                    landmark records into the row buffers with LDGSTS copies and prefetches the row's knot / pair records into L1,
//                    both under the observation-pose math; every thread builds its whole 912-B row in shared memory and the row
//                    leaves with ONE TMA bulk store to the caller's row index (one store per 32-row tile with KTK_EVAL_DEVICE_ORDER)
//   k_static_rs_local  the same rows in tangent coordinates (KTK_EVAL_LOCAL), staged in two halves and scattered cooperatively
//   k_newton_rs_fast + k_newton_rs_rev   NewtonRs rows in closed form: static row at the last row time of the iteration + pi'(t_last) (x) d t_last / d theta
//                    (one reverse sweep per evaluation); k_newton_rs: forward mode through the iteration for the rows that path cannot do
//   k_lifting_rs     LiftingRs rows, closed form
//   k_*_split        the same measurements on a split (R3 + SO3) trajectory; k_span_rs_split: NewtonRs / LiftingRs there (forward mode)
//   k_imu_sensor, k_static_rs_sensor, k_span_sensor   columns of the sensors' own parameter blocks (KTK_EVAL_SENSOR_JACOBIANS); k_span_localize*: local rows
//   k_gn_* (gn_device.cuh)   Gauss-Newton / LM step on the rows left in device memory
// Measurement records are sorted once (first evaluation) by their first active knot so that a warp touches one or two
// knot windows; a thread builds its Jacobian row in shared memory and the finished row leaves with one TMA bulk store
// (cp.async.bulk.global.shared::cta) at the CALLER's row index, so rows come out in insertion order without a second pass
// (the short rows of the split IMU kernels and the local camera rows use a cooperative 16-byte scatter instead).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/kontiki_b200.h"
#include "sensor_jac.cuh"
#include "newton_math.cuh"

using namespace kb;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
inline bool is_camera(int kind) { return kind == KTK_STATIC_RS || kind == KTK_NEWTON_RS || kind == KTK_LIFTING_RS; }
inline bool is_span_camera(int kind) { return kind == KTK_NEWTON_RS || kind == KTK_LIFTING_RS; }      // rows carry the whole observation span
#define KTK_CUDA(call)                                                                                   \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(KTK_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));   \
  } while (0)

// error paths fill a row with NaN: keep that loop rolled, it never runs in a healthy evaluation
#define KTK_COLD_LOOP _Pragma("unroll 1")
#ifndef KTK_THREADS
#define KTK_THREADS 32
#endif
#ifndef KTK_GYRO_MINB
#define KTK_GYRO_MINB 1
#endif
#ifndef KTK_ACCEL_MINB
#define KTK_ACCEL_MINB 1
#endif
#ifndef KTK_CAM_THREADS
#define KTK_CAM_THREADS 32
#endif
#ifndef KTK_CAM_MINB
#define KTK_CAM_MINB 1
#endif
constexpr int kThreads = KTK_THREADS;            // measurement rows per CTA (one per thread), IMU and landmark kernels
constexpr int kCamThreads = KTK_CAM_THREADS;     // ... static-RS observation kernel
constexpr int kImuRow = 84, kImuRowStride = 86;     // doubles; stride keeps rows 16-B aligned and off the same banks
constexpr int kAccelRowStride = 110;                 // accelerometer rows: + 24 doubles of scratch behind the row (accel_se3<PARK>)
__host__ __device__ constexpr int imu_stride(int which) { return which == 1 ? kAccelRowStride : kImuRowStride; }
constexpr int kCamRow = 114;                          // packed camera row in global memory: 112 knot-block doubles + d r/d rho (2)
#ifndef KTK_CAM_STRIDE
#define KTK_CAM_STRIDE 92
#endif
// The SE3 camera kernel stages a row in two halves through ONE 92-double buffer per row (the landmark record lands in
// it, the reference-window half is produced in place and scattered, then the observation-window half re-uses it):
// 23 KB per warp => 8 warps of rows per SM (register-limited) AND ~40 KB of L1 left for the pair table.  L1 matters:
// with 4 KB left (8 warps x 28 KB) the kernel took 0.40 ms, with 24 KB (7 x 29 KB) 0.27 ms (profiles/README.md).
#ifndef KTK_WINDOW
#define KTK_WINDOW 0
#endif
constexpr int kWinCap = KTK_WINDOW;   // pair records of the tile window staged in shared memory; 0 (default) = read through L1: staging measured no gain
constexpr int kCamHalf = 56, kCamRowStride = KTK_CAM_STRIDE;
constexpr int kCamStage = 112, kCamSplitStride = 114;  // the split-trajectory kernel still stages the whole row
constexpr int kCamWarpSmem = 32 * kCamRowStride + kWinCap * kPairStride;   // doubles of shared memory per warp: 32 row buffers + pair window
constexpr int kCamSplitWarpSmem = 32 * kCamSplitStride;

// ---- data movement helpers ---------------------------------------------------------------------------------------
// One TMA bulk store shared -> global (issued by ONE lane for a whole warp tile; UBLKCP is a uniform-datapath
// instruction, per-lane issue serialises the warp).
__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, unsigned bytes) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16-byte asynchronous copy global -> shared (LDGSTS)
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Warp-cooperative scatter of the warp's 32 staged rows (STAGE doubles each, STRIDE apart in shared memory) to their
// rows (GROW doubles apart) in global memory: for each row, the 32 lanes move consecutive 16-byte chunks, so every store
// instruction writes one contiguous run.  dst_row < 0 skips a row (ragged tail / nothing to write).
template <int STAGE, int STRIDE, int GROW>
__device__ __forceinline__ void warp_scatter_rows(const double* wbase, double* gJ, int dst_row, int lane) {   // gJ may carry a column offset
  constexpr int kChunks = STAGE / 2;     // 16-byte chunks per row
  static_assert(kChunks <= 64, "at most two passes of 32 lanes");
  const bool full = __all_sync(0xffffffffu, dst_row >= 0);
  if (full) {
#pragma unroll 8
    for (int rr = 0; rr < 32; ++rr) {
      const int d = __shfl_sync(0xffffffffu, dst_row, rr);
      const double2* src = reinterpret_cast<const double2*>(wbase + rr * STRIDE);
      double2* dst = reinterpret_cast<double2*>(gJ + (size_t)(unsigned)d * GROW);
      if (kChunks >= 32 || lane < kChunks) dst[lane] = src[lane];
      if (kChunks > 32 && lane < kChunks - 32) dst[32 + lane] = src[32 + lane];
    }
  } else {
    for (int rr = 0; rr < 32; ++rr) {
      const int d = __shfl_sync(0xffffffffu, dst_row, rr);
      if (d < 0) continue;
      const double2* src = reinterpret_cast<const double2*>(wbase + rr * STRIDE);
      double2* dst = reinterpret_cast<double2*>(gJ + (size_t)d * GROW);
      for (int c = lane; c < kChunks; c += 32) dst[c] = src[c];
    }
  }
}
// ... and the gather of 32 records (REC doubles each) from global memory into the row buffers at offset OFF (LDGSTS).
template <int REC, int STRIDE, int OFF>
__device__ __forceinline__ void warp_gather_records(double* wbase, const double* recs, int ridx, int lane) {
  constexpr int kChunks = REC / 2;
  const bool full = __all_sync(0xffffffffu, ridx >= 0);
  if (full) {
#pragma unroll 8
    for (int rr = 0; rr < 32; ++rr) {
      const int ri = __shfl_sync(0xffffffffu, ridx, rr);
      const double2* src = reinterpret_cast<const double2*>(recs + (size_t)ri * REC);
      double2* dst = reinterpret_cast<double2*>(wbase + rr * STRIDE + OFF);
      if (kChunks >= 32 || lane < kChunks) cp_async16(dst + lane, src + lane);
      if (kChunks > 32 && lane < kChunks - 32) cp_async16(dst + 32 + lane, src + 32 + lane);
    }
  } else {
    for (int rr = 0; rr < 32; ++rr) {
      const int ri = __shfl_sync(0xffffffffu, ridx, rr);
      if (ri < 0) continue;
      const double2* src = reinterpret_cast<const double2*>(recs + (size_t)ri * REC);
      double2* dst = reinterpret_cast<double2*>(wbase + rr * STRIDE + OFF);
      for (int c = lane; c < kChunks; c += 32) cp_async16(dst + c, src + c);
    }
  }
}

// K0, fused with the knot packing: reads the caller's n x 7 knots, writes the 64-B knot records (thread dir == 14 of pair p packs knot p,
// pair 1 also knot 0) and the pair records.  One launch instead of two: the step is a chain of short kernels and each link costs ~3 us.
__global__ void k_pair_prepass(const double* __restrict__ k7, int n_knots, double* __restrict__ k8, double* __restrict__ pairs, int* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && err) *err = 0;      // the first kernel of an evaluation clears the status word (every writer runs in a later kernel): no memset node
  const int p = 1 + i / 15, dir = i % 15;
  if (p >= n_knots) return;
  pair_prepass_item<7>(k7, p, dir, pairs);
  if (dir == 14) {
    for (int q = (p == 1 ? 0 : p); q <= p; ++q) {
#pragma unroll
      for (int c = 0; c < 7; ++c) k8[(size_t)q * kKnotStride + c] = k7[(size_t)q * 7 + c];
      k8[(size_t)q * kKnotStride + 7] = 0.0;
    }
  }
}

__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
// One warp = one tile of 32 consecutive sorted rows; CTAs are small (1-2 warps) and one-shot: the hardware CTA scheduler
// balances them better than a persistent loop did (measured: profiles/README.md "experiments").
__device__ __forceinline__ int warp_tile() { return blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); }

struct ImuArgs {
  SplineConst sp; ImuConst imu;
  const double* knots; const double* pairs;
  const double* t; const double* y; const double* w; const int* perm;
  int n; uint32_t flags;
  int ahead;                           // tiles resident on the chip: distance of the L2 prefetch of the row inputs (0 = off)
  double* r; double* J; int* i0; int* err;
};
struct ImuIn { double t, y0, y1, y2, y3, w; int perm; };
__device__ __forceinline__ ImuIn imu_load_q(const ImuArgs& a, int i) {      // OrientationMeasurement: y = q, 4 doubles per row
  ImuIn in; in.perm = -1; in.t = 0; in.y0 = in.y1 = in.y2 = in.y3 = 0; in.w = 0;
  if (i < a.n) {
    const double2 q0 = reinterpret_cast<const double2*>(a.y)[2 * (size_t)i], q1 = reinterpret_cast<const double2*>(a.y)[2 * (size_t)i + 1];
    in.t = a.t[i]; in.y0 = q0.x; in.y1 = q0.y; in.y2 = q1.x; in.y3 = q1.y; in.w = a.w[i];
    in.perm = (a.flags & KTK_EVAL_DEVICE_ORDER) ? i : a.perm[i];
  }
  return in;
}
__device__ __forceinline__ ImuIn imu_load(const ImuArgs& a, int i) {
  ImuIn in; in.perm = -1; in.t = 0; in.y0 = in.y1 = in.y2 = in.y3 = 0; in.w = 0;
  if (i < a.n) {
    in.t = a.t[i]; in.y0 = a.y[3 * (size_t)i]; in.y1 = a.y[3 * (size_t)i + 1]; in.y2 = a.y[3 * (size_t)i + 2]; in.w = a.w[i];
    in.perm = (a.flags & KTK_EVAL_DEVICE_ORDER) ? i : a.perm[i];      // destination row: device order or the caller's insertion order
  }
  return in;
}

// WHICH: 0 gyroscope, 1 accelerometer, 2 PositionMeasurement (3 residuals, y[3], rows [4][3][7]); 3 OrientationMeasurement (ONE residual,
// y = q (x,y,z,w), rows [4][1][7])
template <int WHICH>
__device__ __forceinline__ void imu_tile(const ImuArgs& a, int tile, double* wbase) {
  constexpr int NR = WHICH == 3 ? 1 : 3, ROW = NR * 28, LROW = NR * 24, STRIDE = imu_stride(WHICH);
  const int lane = threadIdx.x & 31;
  double* row = wbase + lane * STRIDE;
  if (tile * 32 >= a.n) return;
  const bool wantJ = a.J && (a.flags & KTK_EVAL_JACOBIANS);
  const bool local = (a.flags & KTK_EVAL_LOCAL) != 0;
  const int i = tile * 32 + lane;
  if (a.ahead > 0) {      // same as k_static_rs: a one-shot CTA meets its inputs cold; pull those of the tile that starts when this one ends into L2
    const long long i2 = 32ll * ((long long)tile + a.ahead);
    if (i2 + 32 <= a.n) {
      constexpr int NYP = WHICH == 3 ? 4 : 3;
      if (lane < 2) prefetch_l2(a.t + i2 + 16 * lane);
      else if (lane < 4) prefetch_l2(a.w + i2 + 16 * (lane - 2));
      else if (lane == 4) prefetch_l2(a.perm + i2);
      else if (lane < 5 + 2 * NYP) prefetch_l2(a.y + NYP * i2 + 16 * (lane - 5));
    }
  }
  const ImuIn cur = WHICH == 3 ? imu_load_q(a, i) : imu_load(a, i);
  if (cur.perm >= 0) {
    double y[4] = {cur.y0, cur.y1, cur.y2, cur.y3};
    if (WHICH != 3) { y[0] -= a.imu.bias[0]; y[1] -= a.imu.bias[1]; y[2] -= a.imu.bias[2]; }   // r = w (y - (model + bias))
    double r[3];
    int i0 = -1;
    const int st = imu_row<WHICH == 1>(WHICH, a.sp, a.imu, a.knots, a.pairs, cur.t, y, cur.w, r, row, &i0, row + kImuRowStride);
    if (st != 0) {
      atomicMin(a.err, st);
      r[0] = r[1] = r[2] = nan("");
      KTK_COLD_LOOP for (int c = 0; c < ROW; ++c) row[c] = nan("");
    }
    else if (local) localize_se3_blocks<NR>(row, 4, a.knots + (size_t)i0 * kKnotStride);
    const size_t dst = (size_t)cur.perm;
    if (a.r) {
#pragma unroll
      for (int c = 0; c < NR; ++c) a.r[NR * dst + c] = r[c];
    }
    if (a.i0) a.i0[dst] = i0;
  }
#ifndef KTK_IMU_TMA
#define KTK_IMU_TMA 1
#endif
#if KTK_IMU_TMA      // every finished row leaves with one TMA bulk store (672 / 576 / 224 / 192 B), like the camera rows
  fence_async_smem();
  __syncwarp();
  if (wantJ && cur.perm >= 0) {
    const int len = local ? LROW : ROW;
    bulk_store(a.J + (size_t)cur.perm * len, row, (unsigned)(len * 8));
    bulk_store_wait_read();
  }
#else
  __syncwarp();
  if (wantJ) {
    if (local) warp_scatter_rows<LROW, STRIDE, LROW>(wbase, a.J, cur.perm, lane);
    else warp_scatter_rows<ROW, STRIDE, ROW>(wbase, a.J, cur.perm, lane);
  }
#endif
}
template <int WHICH>
__global__ void __launch_bounds__(kThreads, WHICH == 0 ? KTK_GYRO_MINB : KTK_ACCEL_MINB) k_imu(const ImuArgs a) {
  extern __shared__ __align__(16) double smem[];
  imu_tile<WHICH>(a, warp_tile(), smem + (size_t)(threadIdx.x >> 5) * 32 * imu_stride(WHICH));
}

struct RefArgs {
  SplineConst sp; CameraConst cam;
  const double* knots; const double* pairs; const double* rho;
  const double* ref_uv; const double* ref_t0; const int* seg_start; const int* seg_n; const int* lm;
  int n; double* recs; int* err;
};

// Landmark-reference records come out in record order: a warp's 32 records are one contiguous 32 x 736 B block,
// written with a single TMA bulk store issued by lane 0.
__device__ __forceinline__ void landmark_tile(const RefArgs& a, int tile, double* wbase) {
  const int lane = threadIdx.x & 31;
  double* row = wbase + lane * kRefStride;
  const int base = tile * 32;
  const int i = base + lane;
  if (base >= a.n) return;
  if (i < a.n) {
    const double ruv[2] = {a.ref_uv[2 * (size_t)i], a.ref_uv[2 * (size_t)i + 1]};
    const int st = landmark_ref_row(a.sp, a.cam, a.knots, a.pairs, ruv, a.ref_t0[i], a.seg_start[i], a.seg_n[i], a.rho[a.lm[i]], row);
    if (st != 0) {
      atomicMin(a.err, st);
      KTK_COLD_LOOP for (int c = 0; c < kRefStride; ++c) row[c] = nan("");
      row[7] = -1.0;
    }
  }
  fence_async_smem();
  __syncwarp();
  if (lane == 0) {
    bulk_store(a.recs + (size_t)base * kRefStride, wbase, (unsigned)(min(32, a.n - base) * kRefStride * 8));
    bulk_store_wait_read();
  }
}
__global__ void __launch_bounds__(kThreads) k_landmark_ref(const RefArgs a) {
  extern __shared__ __align__(16) double smem[];
  landmark_tile(a, warp_tile(), smem + (size_t)(threadIdx.x >> 5) * 32 * kRefStride);
}

// The "short" kernels of one evaluation -- every IMU-like group (gyroscope / accelerometer / position / orientation rows) and every camera
// group's landmark-record table -- in ONE launch: each is about one wave of one-warp CTAs on its own (1.32 waves on H1: two tile latencies
// where 1.3 would do, three times per evaluation).  Segments are contiguous ranges of blockIdx, so co-resident CTAs mostly run the same
// code path (running them as concurrent kernels lost to instruction-cache thrash, profiles/r1g_kernel_experiments.md).
constexpr int kShortMaxImu = 4, kShortMaxRef = 2;
struct ShortBatch {
  int n_imu, n_ref;
  int first[kShortMaxImu + kShortMaxRef + 1];      // first CTA of every segment (IMU segments, then landmark segments), then the total
  int which[kShortMaxImu];
  ImuArgs imu[kShortMaxImu];
  RefArgs ref[kShortMaxRef];
};
__global__ void __launch_bounds__(32) k_short_batch(const __grid_constant__ ShortBatch b) {
  extern __shared__ __align__(16) double smem[];
  const int cta = blockIdx.x;
#pragma unroll 1
  for (int k = 0; k < b.n_imu; ++k)
    if (cta < b.first[k + 1]) {
      const int tile = cta - b.first[k];
      if (b.which[k] == 0) imu_tile<0>(b.imu[k], tile, smem);
      else if (b.which[k] == 1) imu_tile<1>(b.imu[k], tile, smem);
      else if (b.which[k] == 2) imu_tile<2>(b.imu[k], tile, smem);
      else imu_tile<3>(b.imu[k], tile, smem);
      return;
    }
#pragma unroll 1
  for (int k = 0; k < b.n_ref; ++k)
    if (cta < b.first[b.n_imu + k + 1]) { landmark_tile(b.ref[k], cta - b.first[b.n_imu + k], smem); return; }
}

struct CamArgs {
  SplineConst sp; CameraConst cam;
  const double* knots; const double* pairs; const double* recs;
  const double* obs_uv; const double* obs_t0; const double* ref_t0; const int* ref_idx; const double* w; const double* huber;
  const int* perm;
  const int* io; const double* uo;     // first knot / interpolation amount of the observation evaluation, located once at upload time
  int n; uint32_t flags;
  int ahead;                           // k_static_rs: tiles resident on the chip (0 = no input prefetch)
  double* r; double* J; int* i0r; int* i0o; int* err;
};
struct CamIn { double u, v, uo, w, huber; int io, ridx, perm; };
__device__ __forceinline__ CamIn cam_load(const CamArgs& a, int i) {
  CamIn in; in.perm = -1; in.ridx = -1; in.io = -1; in.u = in.v = in.uo = in.w = in.huber = 0;
  if (i < a.n) {
    in.u = a.obs_uv[2 * (size_t)i]; in.v = a.obs_uv[2 * (size_t)i + 1]; in.io = a.io[i]; in.uo = a.uo[i]; in.w = a.w[i];
    in.huber = (a.flags & KTK_EVAL_ROBUST) ? a.huber[i] : 0.0; in.ridx = a.ref_idx[i];
    in.perm = (a.flags & KTK_EVAL_DEVICE_ORDER) ? i : a.perm[i];
  }
  return in;
}

// KTK_EVAL_LOCAL rows (2 x 6 blocks, 98 doubles): staged in two halves through one 92-double buffer, cooperative scatter.
__global__ void __launch_bounds__(kCamThreads, KTK_CAM_MINB) k_static_rs_local(const CamArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31;
  double* wbase = smem + (size_t)(threadIdx.x >> 5) * kCamWarpSmem;
  double* row = wbase + lane * kCamRowStride;
  const int tile = warp_tile();
  if (tile * 32 >= a.n) return;
  const bool wantJ = a.J && (a.flags & KTK_EVAL_JACOBIANS);
  const bool local = (a.flags & KTK_EVAL_LOCAL) != 0;
  const int grow = local ? 98 : kCamRow;        // doubles per row in global memory: [ref half | obs half | d r/d rho (2)]
  const CamIn cur = cam_load(a, tile * 32 + lane);
  const double ouv[2] = {cur.u, cur.v};
  const bool live = cur.perm >= 0 && cur.ridx >= 0;
  const double* pairs = a.pairs;
  int wmin = 0, wmax = -1;
#if KTK_WINDOW > 0
  // Stage the tile's window of the pair table: rows are sorted by first knot, so the 32 rows share 3 + (i0max - i0min)
  // consecutive 832-B records.  A one-warp CTA meets a cold L1 and the row scatter streams through it, so without this
  // every sector of the window is fetched from L2 several times per tile (ncu: 18 % L1 hit rate on the loads).
  {
    const int ig = live ? cur.io : 0;
    wmin = __reduce_min_sync(0xffffffffu, live ? ig : 0x7fffffff);
    wmax = __reduce_max_sync(0xffffffffu, live ? ig : (int)0x80000000);
    const int nrec = wmax - wmin + 3;
    if (wmin <= wmax && nrec <= kWinCap && wmin >= 0 && wmax + 3 < a.sp.n_knots) {
      const double2* src = reinterpret_cast<const double2*>(a.pairs + (size_t)(wmin + 1) * kPairStride);
      double2* dst = reinterpret_cast<double2*>(wbase + 32 * kCamRowStride);
      for (int c = lane; c < nrec * (kPairStride / 2); c += 32) cp_async16(dst + c, src + c);
    } else {
      wmax = wmin - 1;      // nothing staged: every lane reads the global table
    }
    cp_async_commit();
  }
#endif
  // gather the 32 landmark records of this tile into the row buffers (cooperative 16-B LDGSTS: every instruction moves
  // one contiguous 512-B run), in flight during the observation-pose evaluation below
  warp_gather_records<kRefStride, kCamRowStride, kRefInRow>(wbase, a.recs, cur.ridx, lane);
  cp_async_commit();
  ObsForward f; f.status = kStatusRange; f.io = -1;
  if (live && cur.io >= 0) { f.status = 0; f.io = cur.io; f.bo = cumulative_basis(cur.uo, a.sp.dt); }
  // a row whose exact first knot (segment arithmetic) lies in the staged window reads its pair records from shared memory
  if (f.status == 0 && f.io >= wmin && f.io <= wmax) pairs = wbase + 32 * kCamRowStride - (size_t)(wmin + 1) * kPairStride;
  cp_async_wait_group<1>();      // the window has landed; the record gather may still be in flight
  __syncwarp();
  static_rs_row_pose(a.knots, pairs, f);
  cp_async_wait_all();
  __syncwarp();
  ObsAdjoint adj;
  int st = 0;
  if (cur.perm >= 0) {
    double r[2], jrho[2];
    int ir = -1, io = -1;
    st = static_rs_row_ref_half(a.cam, f, row + kRefInRow, ouv, cur.w, cur.huber, r, row, jrho, &ir, &io, adj);
    if (st != 0) {
      atomicMin(a.err, st);
      r[0] = r[1] = jrho[0] = jrho[1] = nan(""); ir = io = -1;
      KTK_COLD_LOOP for (int c = 0; c < kCamHalf; ++c) row[c] = nan("");
    } else if (local) localize_se3_blocks<2>(row, 4, a.knots + (size_t)ir * kKnotStride);
    const size_t dst = (size_t)cur.perm;
    if (a.r) { a.r[2 * dst] = r[0]; a.r[2 * dst + 1] = r[1]; }
    if (a.i0r) a.i0r[dst] = ir;
    if (a.i0o) a.i0o[dst] = io;
    if (wantJ) *reinterpret_cast<double2*>(a.J + dst * grow + (grow - 2)) = make_double2(jrho[0], jrho[1]);
  }
  __syncwarp();
  if (wantJ) {                                                                                               // reference-window half
    if (local) warp_scatter_rows<48, kCamRowStride, 98>(wbase, a.J, cur.perm, lane);
    else warp_scatter_rows<kCamHalf, kCamRowStride, kCamRow>(wbase, a.J, cur.perm, lane);
  }
  __syncwarp();
  if (cur.perm >= 0) {
    if (st == 0) { static_rs_row_obs_half(a.knots, pairs, f, adj, row); if (local) localize_se3_blocks<2>(row, 4, a.knots + (size_t)f.io * kKnotStride); }
    else for (int c = 0; c < kCamHalf; ++c) row[c] = nan("");
  }
  __syncwarp();
  if (wantJ) {                                                                                               // observation-window half
    if (local) warp_scatter_rows<48, kCamRowStride, 98>(wbase, a.J + 48, cur.perm, lane);
    else warp_scatter_rows<kCamHalf, kCamRowStride, kCamRow>(wbase, a.J + kCamHalf, cur.perm, lane);
  }
}

// The static-RS kernel for ambient rows.  Every thread builds its whole 114-double row in shared memory (the landmark record is
// gathered at offset 22 so that the in-place rewrite of the reference-window blocks never overtakes its reads) and the row leaves
// with ONE TMA bulk store (cp.async.bulk.global.shared::cta, 912 B) to the caller's row index: no scatter loop, no second
// staging pass.  With KTK_EVAL_DEVICE_ORDER row k of the output is the k-th row in DEVICE order (sorted by first knot;
// ktk_get_row_order gives the insertion index): the warp's 32 rows are one contiguous 29-KB block and lane 0 issues a single
// bulk store for the tile.
constexpr int kCamDevStride = 114, kRefInRowDev = 22;      // record at 22..113: block k read at 30 + 21 k, written at 14 k
#ifdef KTK_PHASE_TIMING      // experimental builds (tools/phase_timing.py): clock64 sums per phase of a tile, lane 0 of every warp
__device__ unsigned long long g_phase[8];
#define KTK_PHASE(k) do { if (lane == 0) { const long long t_ = clock64(); atomicAdd(&g_phase[k], (unsigned long long)(t_ - tphase)); tphase = t_; } } while (0)
#else
#define KTK_PHASE(k) do { } while (0)
#endif
__global__ void __launch_bounds__(kCamThreads, KTK_CAM_MINB) k_static_rs(const CamArgs a) {
  extern __shared__ __align__(16) double smem[];
#ifdef KTK_PHASE_TIMING
  long long tphase = clock64();
#endif
  const int lane = threadIdx.x & 31;
  double* wbase = smem + (size_t)(threadIdx.x >> 5) * 32 * kCamDevStride;
  double* row = wbase + lane * kCamDevStride;
  const int tile = warp_tile();
  if (tile * 32 >= a.n) return;
  const bool wantJ = a.J && (a.flags & KTK_EVAL_JACOBIANS);
  const int i = tile * 32 + lane;
  // A one-shot CTA meets its row inputs cold (one exposed DRAM round trip, 10 % of the stall samples): pull the inputs of the tile
  // `ahead` tiles further on -- about what is resident on the chip, i.e. a CTA that starts when this one ends -- into L2 now.
  if (a.ahead > 0) {
    const long long i2 = 32ll * ((long long)tile + a.ahead);
    if (i2 + 32 <= a.n) {
      if (lane < 4) prefetch_l2(a.obs_uv + 2 * i2 + 16 * lane);       // 512 B of (u, v)
      else if (lane < 6) prefetch_l2(a.uo + i2 + 16 * (lane - 4));    // 256 B each below
      else if (lane < 8) prefetch_l2(a.w + i2 + 16 * (lane - 6));
      else if (lane < 10) prefetch_l2(a.huber + i2 + 16 * (lane - 8));
      else if (lane == 10) prefetch_l2(a.io + i2);                     // 128 B each
      else if (lane == 11) prefetch_l2(a.ref_idx + i2);
      else if (lane == 12) prefetch_l2(a.perm + i2);
    }
  }
  CamIn cur = cam_load(a, i);
  KTK_PHASE(0);      // prefetch issue + input loads issued
#ifdef KTK_ABL_FIXED_IO      // timing ablations only (tools/ablate.sh): wrong results on purpose
  if (cur.io >= 0) cur.io = 100;
#endif
#ifdef KTK_ABL_NOGATHER
  if (cur.ridx >= 0) cur.ridx = 0;
#endif
  const double ouv[2] = {cur.u, cur.v};
#ifndef KTK_L1_WINDOW
#define KTK_L1_WINDOW 1
#endif
#if KTK_L1_WINDOW == 1      // pull the row's knot record and three pair records (2.5 KB, mostly shared by the warp) into L1 while the gather is issued:
  // the forward sweep's first use of them was an exposed L2 round trip (6 % of the stall samples).  0.2115 -> 0.2078 ms (profiles/r2w)
  if (cur.io >= 0) {
    const char* w0 = reinterpret_cast<const char*>(a.pairs + (size_t)(cur.io + 1) * kPairStride);
    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.knots + (size_t)cur.io * kKnotStride));
#pragma unroll
    for (int k = 0; k < 3 * kPairStride * 8; k += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(w0 + k));
  }
#elif KTK_L1_WINDOW == 3    // variant: one line per lane of the window of the warp's first knot (4 pair records + 2 knot records)
  {
    const int wmin = __reduce_min_sync(0xffffffffu, cur.io >= 0 ? cur.io : 0x7fffffff);
    if (wmin != 0x7fffffff) {
      const char* w0 = reinterpret_cast<const char*>(a.pairs + (size_t)(wmin + 1) * kPairStride);
      const int last = a.sp.n_knots - 1;
      const int nb = (min(wmin + 4, last) - wmin) * kPairStride * 8;
      if (lane * 128 < nb) asm volatile("prefetch.global.L1 [%0];" ::"l"(w0 + lane * 128));
      if (lane == 31) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.knots + (size_t)wmin * kKnotStride));
    }
  }
#endif
  warp_gather_records<kRefStride, kCamDevStride, kRefInRowDev>(wbase, a.recs, cur.ridx, lane);
  KTK_PHASE(1);      // waited for the inputs (ridx), gather issued
  ObsForward f; f.status = kStatusRange; f.io = -1;
  if (cur.perm >= 0 && cur.ridx >= 0 && cur.io >= 0) {
    f.status = 0; f.io = cur.io; f.bo = cumulative_basis(cur.uo, a.sp.dt);
    static_rs_row_pose(a.knots, a.pairs, f);
  }
  KTK_PHASE(2);      // observation pose (forward sweep)
  cp_async_wait_all();
  __syncwarp();
  KTK_PHASE(3);      // waited for the landmark records
  if (cur.perm >= 0) {
    double r[2], jrho[2];
    int ir = -1, io = -1;
    ObsAdjoint adj;
    const int st = static_rs_row_ref_half(a.cam, f, row + kRefInRowDev, ouv, cur.w, cur.huber, r, row, jrho, &ir, &io, adj);
    KTK_PHASE(4);    // projection + reference-window half
#ifdef KTK_ST_EARLY      // experiment: the finished reference-window half leaves while the observation-window half is computed
    if (st == 0 && wantJ && !(a.flags & KTK_EVAL_DEVICE_ORDER)) { fence_async_smem(); bulk_store(a.J + (size_t)cur.perm * kCamRow, row, (unsigned)(kCamHalf * 8)); }
#endif
    // Everything the first half produced leaves BEFORE the reverse sweep: d r / d rho (the record under row[112..113] has been consumed), the
    // residual and the two indices.  Stored after it they were live across the sweep, and two of them were spilled: a reload from local
    // memory misses the 28 KB of L1 this kernel leaves and is an L2 round trip (ncu, profiles/r2m: ~4 % of the samples on its first use).
    if (st != 0) {
      atomicMin(a.err, st);
      r[0] = r[1] = nan(""); ir = io = -1;
      KTK_COLD_LOOP for (int c = 0; c < kCamRow; ++c) row[c] = nan("");
    } else { row[112] = jrho[0]; row[113] = jrho[1]; }
    const size_t dst = (size_t)cur.perm;                 // == i in device order
    if (a.r) { a.r[2 * dst] = r[0]; a.r[2 * dst + 1] = r[1]; }
    if (a.i0r) a.i0r[dst] = ir;
    if (a.i0o) a.i0o[dst] = io;
    if (st == 0) {
#ifndef KTK_ABL_NOBACK
      static_rs_row_obs_half(a.knots, a.pairs, f, adj, row + kCamHalf);
#else
      for (int c = 0; c < 18; ++c) row[kCamHalf + c] = adj.Gp.a[c % 6] + adj.GpR.a[c % 6] * adj.Gth.a[c % 6];
#endif
    }
  }
  KTK_PHASE(5);      // observation-window half (reverse sweep) + r / index stores
  fence_async_smem();
  __syncwarp();
  if (!wantJ) return;
#ifdef KTK_ABL_NOSTORE
  if (cur.u != -12345.678) return;
#endif
  if (a.flags & KTK_EVAL_DEVICE_ORDER) {
    if (lane == 0) {
      bulk_store(a.J + (size_t)tile * 32 * kCamRow, wbase, (unsigned)(min(32, a.n - tile * 32) * kCamRow * 8));
      bulk_store_wait_read();
    }
  }
#if defined(KTK_ST_STG)        // experiment: cooperative 16-byte stores instead of TMA
  else { warp_scatter_rows<kCamRow, kCamDevStride, kCamRow>(wbase, a.J, cur.perm, lane); }
#elif defined(KTK_ST_EARLY)
  else if (cur.perm >= 0) {
    if (row[0] == row[0]) bulk_store(a.J + (size_t)cur.perm * kCamRow + kCamHalf, row + kCamHalf, (unsigned)((kCamRow - kCamHalf) * 8));
    else bulk_store(a.J + (size_t)cur.perm * kCamRow, row, (unsigned)(kCamRow * 8));      // NaN row (error path): nothing was sent early
    bulk_store_wait_read();
  }
#else
  else if (cur.perm >= 0) {                            // caller order: one 912-B bulk store per row, to the row's insertion index
#ifdef KTK_ST_L2DST
    bulk_store(a.J + (size_t)(cur.perm & 32767) * kCamRow, row, (unsigned)(kCamRow * 8));
#else
    bulk_store(a.J + (size_t)cur.perm * kCamRow, row, (unsigned)(kCamRow * 8));
#endif
    KTK_PHASE(6);    // store issue
#ifndef KTK_ST_NOWAIT
    bulk_store_wait_read();
#endif
    KTK_PHASE(7);    // waited for the TMA engine to read the rows
  }
#endif
}

// =====================================================================================================================
// Split trajectory (R3 + SO3) kernels: same tiling and data movement as above, rows of 48 / 84 / 114 doubles.
// =====================================================================================================================
constexpr int kGyroSplitRow = 48, kGyroSplitStride = 50;
constexpr int kAccelSplitRow = 84, kAccelSplitStride = 86;
constexpr int kPosSplitRow = 36, kPosSplitStride = 38;       // PositionMeasurement on a split trajectory: [4 R3 knots][3][3]
constexpr int kOriSplitRow = 16, kOriSplitStride = 18;       // OrientationMeasurement on a split trajectory: [4 SO3 knots][1][4]

__global__ void k_pack_vecs(const double* __restrict__ v3, int n, double* __restrict__ v4, int* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && err) *err = 0;
  if (i >= n * kVecStride) return;
  const int k = i / kVecStride, c = i % kVecStride;
  v4[i] = c < 3 ? v3[(size_t)k * 3 + c] : 0.0;
}
__global__ void k_so3_pair_prepass(const double* __restrict__ quats, int n, double* __restrict__ pairs, int* err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = 1 + i / 9, dir = i % 9;
  if (p >= n) return;
  const int st = so3_pair_prepass_item(quats, p, dir, pairs);
  if (st != 0) atomicMin(err, st);
}

struct ImuSplitArgs {
  SplitConst sp; ImuConst imu;
  const double* vecs; const double* quats; const double* pairs;
  const double* t; const double* y; const double* w; const int* perm;
  int n; uint32_t flags;
  double* r; double* J; int* i0_r3; int* i0_so3; int* err;
};
template <int WHICH>
__global__ void __launch_bounds__(kThreads) k_imu_split(const ImuSplitArgs a) {
  constexpr int ROW = WHICH == 0 ? kGyroSplitRow : (WHICH == 1 ? kAccelSplitRow : (WHICH == 2 ? kPosSplitRow : kOriSplitRow));
  constexpr int STRIDE = WHICH == 0 ? kGyroSplitStride : (WHICH == 1 ? kAccelSplitStride : (WHICH == 2 ? kPosSplitStride : kOriSplitStride));
  constexpr int NR = WHICH == 3 ? 1 : 3, NY = WHICH == 3 ? 4 : 3;
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31;
  double* wbase = smem + (size_t)(threadIdx.x >> 5) * 32 * STRIDE;
  double* row = wbase + lane * STRIDE;
  const int tile = warp_tile();
  if (tile * 32 >= a.n) return;
  const bool wantJ = a.J && (a.flags & KTK_EVAL_JACOBIANS);
  const bool local = (a.flags & KTK_EVAL_LOCAL) != 0;
  const int i = tile * 32 + lane;
  int perm = -1;
  if (i < a.n) {
    perm = (a.flags & KTK_EVAL_DEVICE_ORDER) ? i : a.perm[i];
    double y[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int c = 0; c < NY; ++c) y[c] = a.y[NY * (size_t)i + c];
    if (WHICH != 3) { y[0] -= a.imu.bias[0]; y[1] -= a.imu.bias[1]; y[2] -= a.imu.bias[2]; }
    double r[3];
    int ia = -1, ib = -1;
    const int st = imu_row_split(WHICH, a.sp, a.imu, a.vecs, a.quats, a.pairs, a.t[i], y, a.w[i], r, row, &ia, &ib);
    if (st != 0) {
      atomicMin(a.err, st);
      r[0] = r[1] = r[2] = nan(""); ia = ib = -1;
      KTK_COLD_LOOP for (int c = 0; c < ROW; ++c) row[c] = nan("");
    } else if (local && WHICH != 2) localize_so3_blocks<NR>(row + (WHICH == 1 ? 36 : 0), 4, a.quats + (size_t)ib * kQuatStride);
    const size_t dst = (size_t)perm;
    if (a.r) {
#pragma unroll
      for (int c = 0; c < NR; ++c) a.r[NR * dst + c] = r[c];
    }
    if (a.i0_r3) a.i0_r3[dst] = ia;
    if (a.i0_so3) a.i0_so3[dst] = ib;
  }
  // (cooperative 16-byte scatter: the split rows are short -- 384 / 672 B -- and one TMA bulk store per row measured slower here:
  //  C5 gyroscope 0.0517 -> 0.0531 ms, accelerometer 0.0699 -> 0.0756 ms, profiles/r2y)
  __syncwarp();
  if (wantJ) {
    if (local && WHICH != 2) warp_scatter_rows<ROW - 4 * NR, STRIDE, ROW - 4 * NR>(wbase, a.J, perm, lane);      // the four SO3 blocks shrink from NR x 4 to NR x 3
    else warp_scatter_rows<ROW, STRIDE, ROW>(wbase, a.J, perm, lane);
  }
}

struct RefSplitArgs {
  SplitConst sp; CameraConst cam;
  const double* vecs; const double* quats; const double* pairs; const double* rho;
  const double* ref_uv; const double* ref_t0; const int* r3_start; const int* r3_n; const int* so3_start; const int* so3_n; const int* lm;
  int n; double* recs; int* err;
};
__global__ void __launch_bounds__(kThreads) k_landmark_ref_split(const RefSplitArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31;
  double* wbase = smem + (size_t)(threadIdx.x >> 5) * 32 * kRefSplitStride;
  double* row = wbase + lane * kRefSplitStride;
  const int base = (blockIdx.x * kThreads + threadIdx.x) & ~31;
  const int i = base + lane;
  if (base >= a.n) return;
  if (i < a.n) {
    const double ruv[2] = {a.ref_uv[2 * (size_t)i], a.ref_uv[2 * (size_t)i + 1]};
    const int st = landmark_ref_row_split(a.sp, a.cam, a.vecs, a.quats, a.pairs, ruv, a.ref_t0[i], a.r3_start[i], a.r3_n[i], a.so3_start[i], a.so3_n[i],
                                          a.rho[a.lm[i]], row);
    if (st != 0) {
      atomicMin(a.err, st);
      KTK_COLD_LOOP for (int c = 0; c < kRefSplitStride; ++c) row[c] = nan("");
      row[7] = -1.0; row[8] = -1.0;
    }
  }
  fence_async_smem();
  __syncwarp();
  if (lane == 0) {
    bulk_store(a.recs + (size_t)base * kRefSplitStride, wbase, (unsigned)(min(32, a.n - base) * kRefSplitStride * 8));
    bulk_store_wait_read();
  }
}

struct CamSplitArgs {
  SplitConst sp; CameraConst cam;
  const double* vecs; const double* quats; const double* pairs; const double* recs;
  const double* obs_uv; const double* obs_t0; const double* ref_t0; const int* ref_idx; const double* w; const double* huber;
  const int* perm;
  int n; uint32_t flags;
  double* r; double* J; int* idx[4]; int* err;
};
__global__ void __launch_bounds__(kCamThreads, KTK_CAM_MINB) k_static_rs_split(const CamSplitArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31;
  double* wbase = smem + (size_t)(threadIdx.x >> 5) * kCamSplitWarpSmem;
  double* row = wbase + lane * kCamSplitStride;
  const int tile = warp_tile();
  if (tile * 32 >= a.n) return;
  const bool wantJ = a.J && (a.flags & KTK_EVAL_JACOBIANS);
  const bool local = (a.flags & KTK_EVAL_LOCAL) != 0;
  const int grow = local ? 98 : kCamRow;
  const int i = tile * 32 + lane;
  const int perm = i < a.n ? ((a.flags & KTK_EVAL_DEVICE_ORDER) ? i : a.perm[i]) : -1;
  const int myridx = i < a.n ? a.ref_idx[i] : -1;
  warp_gather_records<kRefSplitStride, kCamSplitStride, kRefSplitInRow>(wbase, a.recs, myridx, lane);
  double ouv[2] = {0.0, 0.0};
  ObsForwardSplit f; f.status = kStatusRange; f.ia = f.ib = -1;
  if (perm >= 0 && myridx >= 0) {
    ouv[0] = a.obs_uv[2 * (size_t)i]; ouv[1] = a.obs_uv[2 * (size_t)i + 1];
    static_rs_row_forward_split(a.sp, a.cam, a.quats, a.pairs, ouv, a.obs_t0[i], a.ref_t0[i], f);
  }
  cp_async_wait_all();
  __syncwarp();
  if (perm >= 0) {
    double r[2];
    int idx[4] = {-1, -1, -1, -1};
    const double hub = (a.flags & KTK_EVAL_ROBUST) ? a.huber[i] : 0.0;
    double jrho[2];
    const int st = static_rs_row_finish_split(a.cam, a.vecs, a.quats, a.pairs, f, row + kRefSplitInRow, ouv, a.w[i], hub, r, row, jrho, idx);
    if (st != 0) {
      atomicMin(a.err, st);
      r[0] = r[1] = jrho[0] = jrho[1] = nan(""); idx[0] = idx[1] = idx[2] = idx[3] = -1;
      KTK_COLD_LOOP for (int c = 0; c < kCamStage; ++c) row[c] = nan("");
    } else if (local) {      // [ref R3 24 | ref SO3 32 | obs R3 24 | obs SO3 32] -> [24 | 24 | 24 | 24]
      localize_so3_blocks<2>(row + 24, 4, a.quats + (size_t)idx[2] * kQuatStride);
      for (int c = 0; c < 24; ++c) row[48 + c] = row[56 + c];
      localize_so3_blocks<2>(row + 80, 4, a.quats + (size_t)idx[3] * kQuatStride);
      for (int c = 0; c < 24; ++c) row[72 + c] = row[80 + c];
    }
    const size_t dst = (size_t)perm;
    if (wantJ) {
      if (local) *reinterpret_cast<double2*>(a.J + dst * grow + (grow - 2)) = make_double2(jrho[0], jrho[1]);
      else { row[112] = jrho[0]; row[113] = jrho[1]; }      // ambient rows leave whole (below)
    }
    if (a.r) { a.r[2 * dst] = r[0]; a.r[2 * dst + 1] = r[1]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) if (a.idx[k]) a.idx[k][dst] = idx[k];
  }
  if (!wantJ) return;
  if (local) {
    __syncwarp();
    warp_scatter_rows<96, kCamSplitStride, 98>(wbase, a.J, perm, lane);
    return;
  }
  // ambient rows: the staged 114-double row is the packed row; it leaves with one TMA bulk store like the SE3 kernel's (one per tile in device order)
  fence_async_smem();
  __syncwarp();
  if (a.flags & KTK_EVAL_DEVICE_ORDER) {
    if (lane == 0) bulk_store(a.J + (size_t)tile * 32 * kCamRow, wbase, (unsigned)(min(32, a.n - tile * 32) * kCamRow * 8));
  } else if (perm >= 0) {
    bulk_store(a.J + (size_t)perm * kCamRow, row, (unsigned)(kCamRow * 8));
  }
  bulk_store_wait_read();
}

__global__ void k_traj_eval_se3(SplineConst sp, const double* __restrict__ knots, const double* __restrict__ pairs, int n, const double* __restrict__ t,
                                double* __restrict__ out, int* __restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double o[16];
  const int st = traj_eval_se3(sp, knots, pairs, t[i], o);
  for (int c = 0; c < 16; ++c) out[16 * (size_t)i + c] = st == 0 ? o[c] : nan("");
  status[i] = st;
}
__global__ void k_traj_eval_se3_matrices(SplineConst sp, const double* __restrict__ knots, const double* __restrict__ pairs, int n, const double* __restrict__ t,
                                         double* __restrict__ out, int* __restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double o[48];
  const int st = traj_eval_se3_matrices(sp, knots, pairs, t[i], o);
  for (int c = 0; c < 48; ++c) out[48 * (size_t)i + c] = st == 0 ? o[c] : nan("");
  status[i] = st;
}
__global__ void k_traj_eval_split(SplitConst sp, const double* __restrict__ vecs, const double* __restrict__ quats, const double* __restrict__ pairs, int n,
                                  const double* __restrict__ t, double* __restrict__ out, int* __restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double o[16];
  const int st = traj_eval_split(sp, vecs, quats, pairs, t[i], o);
  for (int c = 0; c < 16; ++c) out[16 * (size_t)i + c] = st == 0 ? o[c] : nan("");
  status[i] = st;
}

// =====================================================================================================================
// Gauss-Newton contraction, matrix-free: products with the packed Jacobian rows that an
// evaluation left in device memory.  Parameter vector (ambient): SE3 [knots 7 n | rho]; split [R3 3 n_r3 | SO3 4 n_so3 | rho].
//   k_j_apply : u[row] = J[row] . v          (gather, no atomics)
//   k_jt_apply: y += J[row]^T u[row]          (scatter with fp64 RED; mode 1: y += J[row]^2 elementwise = diag(J^T J))
// Forming J^T J explicitly would need ~1650 scattered atomics per camera row (the ref x obs cross blocks are unique per
// row); a product needs 57, so the normal equations are applied, not formed, and solved by PCG (kontiki_b200/gn.py).
// =====================================================================================================================
struct RowWindows {            // where the blocks of a packed row live in the parameter vector
  int nwin;                    // number of 4-knot windows
  int j_off[4], width[4], col_off[4], slot[4];     // offset in the row, knot width, first column of that spline, index array
  int nk[4];                                        // knots per window: 4, except the observation span of a Newton-RS row
  int nres, row_len, rho_off_in_row;                // residuals per row, doubles per row, offset of d r/d rho (-1: none)
  long long rho_col0;                               // first column of rho
};
struct ApplyArgs {
  RowWindows w; int n; const double* J; const int* idx[4]; const int* lm;
  const double* v; double* u; double* y; int mode;
  // diag(P^T J^T J P) in LOCAL coordinates: P blocks (width x lwidth, row-major) per knot of each window's spline
  const double* P[4]; int lwidth[4]; long long lcol_off[4]; long long lrho_col0;
};
__global__ void k_j_apply(const ApplyArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const double* Jr = a.J + (size_t)i * a.w.row_len;
  double acc[3] = {0.0, 0.0, 0.0};
  for (int w = 0; w < a.w.nwin; ++w) {
    const int wd = a.w.width[w];
    const double* vv = a.v + a.w.col_off[w] + (size_t)wd * a.idx[a.w.slot[w]][i];
    const double* Jw = Jr + a.w.j_off[w];
    for (int k = 0; k < a.w.nk[w]; ++k)
      for (int r = 0; r < a.w.nres; ++r) {
        double s = 0.0;
        for (int c = 0; c < wd; ++c) s += Jw[(k * a.w.nres + r) * wd + c] * vv[k * wd + c];
        acc[r] += s;
      }
  }
  if (a.w.rho_off_in_row >= 0) {
    const double vr = a.v[a.w.rho_col0 + a.lm[i]];
    for (int r = 0; r < a.w.nres; ++r) acc[r] += Jr[a.w.rho_off_in_row + r] * vr;
  }
  for (int r = 0; r < a.w.nres; ++r) a.u[(size_t)i * a.w.nres + r] = acc[r];
}
// (J P) of one window: t[k][r][c], k = 0..3 knots, r < nres, c < lw
__device__ __forceinline__ void window_JP(const ApplyArgs& a, int w, const double* Jr, int i, double* t) {
  const int wd = a.w.width[w], lw = a.lwidth[w];
  const int k0 = a.idx[a.w.slot[w]][i];
  const double* Jw = Jr + a.w.j_off[w];
  for (int k = 0; k < 4; ++k) {
    const double* Pk = a.P[w] ? a.P[w] + (size_t)(k0 + k) * wd * lw : nullptr;
    for (int r = 0; r < a.w.nres; ++r)
      for (int c = 0; c < lw; ++c) {
        double s = 0.0;
        if (Pk) for (int m = 0; m < wd; ++m) s += Jw[(k * a.w.nres + r) * wd + m] * Pk[m * lw + c];
        else s = Jw[(k * a.w.nres + r) * wd + c];
        t[(k * 3 + r) * 6 + c] = s;
      }
  }
}
__global__ void k_jtj_diag_local(const ApplyArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const double* Jr = a.J + (size_t)i * a.w.row_len;
  double tA[72], tB[72];
  // windows come in (reference, observation) pairs on the same spline: w and w + nwin/2 (camera); a knot that is in both
  // windows is ONE parameter block, its column of J is the sum of the two blocks
  const int npair = a.w.rho_off_in_row >= 0 ? a.w.nwin / 2 : 0;
  for (int w = 0; w < (npair ? npair : a.w.nwin); ++w) {
    const int lw = a.lwidth[w];
    window_JP(a, w, Jr, i, tA);
    const int kA = a.idx[a.w.slot[w]][i];
    int kB = 0x3fffffff;
    if (npair) { window_JP(a, w + npair, Jr, i, tB); kB = a.idx[a.w.slot[w + npair]][i]; }
    for (int k = 0; k < 4; ++k)
      for (int c = 0; c < lw; ++c) {
        double s = 0.0;
        const int kb = kA + k - kB;          // position of this knot in the other window
        for (int r = 0; r < a.w.nres; ++r) {
          double v = tA[(k * 3 + r) * 6 + c];
          if (npair && kb >= 0 && kb < 4) v += tB[(kb * 3 + r) * 6 + c];
          s += v * v;
        }
        atomicAdd(a.y + a.lcol_off[w] + (size_t)lw * (kA + k) + c, s);
      }
    if (npair)
      for (int k = 0; k < 4; ++k) {
        const int ka = kB + k - kA;
        if (ka >= 0 && ka < 4) continue;     // already counted with window A
        for (int c = 0; c < lw; ++c) {
          double s = 0.0;
          for (int r = 0; r < a.w.nres; ++r) { const double v = tB[(k * 3 + r) * 6 + c]; s += v * v; }
          atomicAdd(a.y + a.lcol_off[w] + (size_t)lw * (kB + k) + c, s);
        }
      }
  }
  if (a.w.rho_off_in_row >= 0) {
    double s = 0.0;
    for (int r = 0; r < a.w.nres; ++r) { const double j = Jr[a.w.rho_off_in_row + r]; s += j * j; }
    atomicAdd(a.y + a.lrho_col0 + a.lm[i], s);
  }
}
__global__ void k_jt_apply(const ApplyArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const double* Jr = a.J + (size_t)i * a.w.row_len;
  double ur[3] = {1.0, 1.0, 1.0};
  if (a.mode == 0) for (int r = 0; r < a.w.nres; ++r) ur[r] = a.u[(size_t)i * a.w.nres + r];
  for (int w = 0; w < a.w.nwin; ++w) {
    const int wd = a.w.width[w];
    double* yy = a.y + a.w.col_off[w] + (size_t)wd * a.idx[a.w.slot[w]][i];
    const double* Jw = Jr + a.w.j_off[w];
    for (int k = 0; k < a.w.nk[w]; ++k)
      for (int c = 0; c < wd; ++c) {
        double s = 0.0;
        for (int r = 0; r < a.w.nres; ++r) { const double j = Jw[(k * a.w.nres + r) * wd + c]; s += a.mode == 0 ? j * ur[r] : j * j; }
        atomicAdd(yy + k * wd + c, s);
      }
  }
  if (a.w.rho_off_in_row >= 0) {
    double s = 0.0;
    for (int r = 0; r < a.w.nres; ++r) { const double j = Jr[a.w.rho_off_in_row + r]; s += a.mode == 0 ? j * ur[r] : j * j; }
    atomicAdd(a.y + a.w.rho_col0 + a.lm[i], s);
  }
}

// =====================================================================================================================
// Sensor-block Jacobians (cold path, only with KTK_EVAL_SENSOR_JACOBIANS): one thread per row, rows written straight to
// the caller's row index (3 or 16 doubles).
// =====================================================================================================================
struct ImuSensorArgs {
  int traj, which; SplineConst sp; SplitConst spl; ImuConst imu;
  const double* knots; const double* pairs; const double* vecs; const double* quats; const double* so3pairs;
  const double* t; const double* w; const int* perm; int n; double* Js; int* err; int device_order;
};
__global__ void k_imu_sensor(const ImuSensorArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  double o[3];
  const int st = a.traj == 0 ? imu_time_offset_jac_se3(a.which, a.sp, a.imu, a.knots, a.pairs, a.t[i], a.w[i], o)
                             : imu_time_offset_jac_split(a.which, a.spl, a.imu, a.vecs, a.quats, a.so3pairs, a.t[i], a.w[i], o);
  if (st != 0) { atomicMin(a.err, st); o[0] = o[1] = o[2] = nan(""); }
  double* dst = a.Js + 3 * (size_t)(a.device_order ? i : a.perm[i]);
  dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
}
struct CamSensorArgs {
  int traj; SplitConst spl; const double* vecs; const double* quats; const double* so3pairs;
  SplineConst sp; CameraConst cam; const double* knots; const double* pairs; const double* rho;
  const double* obs_uv; const double* obs_t0; const double* ref_uv; const double* ref_t0; const int* lm; const double* w; const double* huber;
  const int* perm; int n; uint32_t flags; double* Js; int* err;
};
__global__ void k_static_rs_sensor(const CamSensorArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  double o[16];
  const double ouv[2] = {a.obs_uv[2 * (size_t)i], a.obs_uv[2 * (size_t)i + 1]}, ruv[2] = {a.ref_uv[2 * (size_t)i], a.ref_uv[2 * (size_t)i + 1]};
  o[14] = o[15] = 0.0;
  const double hub = (a.flags & KTK_EVAL_ROBUST) ? a.huber[i] : 0.0;
  const int st = a.traj == 0 ? static_rs_sensor_jac_se3(a.sp, a.cam, a.knots, a.pairs, ouv, a.obs_t0[i], ruv, a.ref_t0[i], a.rho[a.lm[i]], a.w[i], hub, o)
                             : static_rs_sensor_jac_split(a.spl, a.cam, a.vecs, a.quats, a.so3pairs, ouv, a.obs_t0[i], ruv, a.ref_t0[i], a.rho[a.lm[i]], a.w[i], hub, o);
  if (st != 0) { atomicMin(a.err, st); for (int c = 0; c < 16; ++c) o[c] = nan(""); }
  double* dst = a.Js + 16 * (size_t)((a.flags & KTK_EVAL_DEVICE_ORDER) ? i : a.perm[i]);
  for (int c = 0; c < 16; ++c) dst[c] = o[c];
}

// =====================================================================================================================
// NewtonRsCameraMeasurement rows (newton_rscamera_measurement.h): the reference's Jacobian is the forward-mode derivative
// through the Newton iteration on the row time, so one thread runs ONE direction of ONE row on a dual number
// (newton_math.cuh); 29 + 7 W directions per row, no staging -- thread (row, dir) owns two doubles of the packed row.
// The landmark side comes from the same k_landmark_ref records as the static measurement.
// =====================================================================================================================
// KTK_EVAL_LOCAL rows of NewtonRs / LiftingRs measurements: the span rows are evaluated in ambient coordinates into a scratch buffer and every knot
// block is then taken through the knot's LocalParameterization (localize_se3_blocks: J_local = J_ambient dPlus/ddelta, uniform_se3_spline_trajectory.h:25-48)
// by one thread per (row, block); local row = [ref 4 x (nres x 6) | obs W x (nres x 6) | tail].  Cold path.
__global__ void k_span_localize(const double* __restrict__ Jamb, const int* __restrict__ i0r, const int* __restrict__ i0o, const double* __restrict__ knots,
                                int n_knots, int n, int W, int nres, int tail, double* __restrict__ Jloc) {
  const int nb = 4 + W + 1;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(tid / nb), b = (int)(tid % nb);
  if (i >= n) return;
  const size_t la = (size_t)(4 + W) * nres * 7 + tail, ll = (size_t)(4 + W) * nres * 6 + tail;
  const double* src = Jamb + (size_t)i * la;
  double* dst = Jloc + (size_t)i * ll;
  if (b == 4 + W) { for (int c = 0; c < tail; ++c) dst[(size_t)(4 + W) * nres * 6 + c] = src[(size_t)(4 + W) * nres * 7 + c]; return; }
  const int first = b < 4 ? i0r[i] : i0o[i];
  int k = first < 0 ? 0 : first + (b < 4 ? b : b - 4);      // blocks past the end of the spline are zero in the ambient row: any knot does
  if (k >= n_knots) k = n_knots - 1;
  double blk[21];
  for (int c = 0; c < nres * 7; ++c) blk[c] = src[(size_t)b * nres * 7 + c];
  if (nres == 3) localize_se3_blocks<3>(blk, 1, knots + (size_t)k * kKnotStride); else localize_se3_blocks<2>(blk, 1, knots + (size_t)k * kKnotStride);
  for (int c = 0; c < nres * 6; ++c) dst[(size_t)b * nres * 6 + c] = blk[c];
}

// ... the same for span rows on a split trajectory: R3 blocks are already local (R^3 is its own tangent space), every SO3 block goes through
// EigenQuaternionParameterization (localize_so3_blocks).  Local row [ref R3 4 x (nres x 3) | ref SO3 4 x (nres x 3) | obs R3 Wa x .. | obs SO3 Wb x .. | tail].
__global__ void k_span_localize_split(const double* __restrict__ Jamb, const int* __restrict__ i0c, const int* __restrict__ i0d, const double* __restrict__ quats,
                                      int n_so3, int n, int Wa, int Wb, int nres, int tail, double* __restrict__ Jloc) {
  const int nb = 8 + Wa + Wb + 1;      // blocks in row order: 4 ref R3, 4 ref SO3, Wa obs R3, Wb obs SO3, tail
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(tid / nb), b = (int)(tid % nb);
  if (i >= n) return;
  const size_t la = (size_t)nres * (28 + 3 * Wa + 4 * Wb) + tail, ll = (size_t)nres * 3 * (8 + Wa + Wb) + tail;
  const double* src = Jamb + (size_t)i * la;
  double* dst = Jloc + (size_t)i * ll;
  if (b == nb - 1) { for (int c = 0; c < tail; ++c) dst[ll - tail + c] = src[la - tail + c]; return; }
  int so3, k, first, soff, doff;      // which spline, knot within its window, first knot, ambient / local offsets of the block
  if (b < 4) { so3 = 0; k = b; first = 0; soff = nres * 3 * k; doff = nres * 3 * k; }
  else if (b < 8) { so3 = 1; k = b - 4; first = i0c[i]; soff = nres * 12 + nres * 4 * k; doff = nres * 12 + nres * 3 * k; }
  else if (b < 8 + Wa) { so3 = 0; k = b - 8; first = 0; soff = nres * 28 + nres * 3 * k; doff = nres * 24 + nres * 3 * k; }
  else { so3 = 1; k = b - 8 - Wa; first = i0d[i]; soff = nres * (28 + 3 * Wa) + nres * 4 * k; doff = nres * (24 + 3 * Wa) + nres * 3 * k; }
  if (!so3) { for (int c = 0; c < nres * 3; ++c) dst[doff + c] = src[soff + c]; return; }
  int kn = first < 0 ? 0 : first + k;
  if (kn >= n_so3) kn = n_so3 - 1;
  double blk[12];
  for (int c = 0; c < nres * 4; ++c) blk[c] = src[soff + c];
  if (nres == 3) localize_so3_blocks<3>(blk, 1, quats + (size_t)kn * kQuatStride); else localize_so3_blocks<2>(blk, 1, quats + (size_t)kn * kQuatStride);
  for (int c = 0; c < nres * 3; ++c) dst[doff + c] = blk[c];
}

// NewtonRs / LiftingRs rows on a SPLIT trajectory (newton_math.cuh "rows on a SPLIT trajectory"): forward mode, one thread per (row, direction);
// every direction writes its own column of the packed row, direction 0 the residual and the four window indices.  Cold path.
struct SpanSplitArgs {
  SplitConst sp; CameraConst cam; const double* vecs; const double* quats; const double* pairs; const double* recs;
  const double* obs_uv; const double* obs_t0; const double* ref_t0; const int* ref_idx; const double* w; const double* huber; const double* vt;
  const int* perm; int n, Wa, Wb, lifting; uint32_t flags;
  double* r; double* J; int* idx[4]; int* err;
};
__global__ void __launch_bounds__(128) k_span_rs_split(const SpanSplitArgs a) {
  const bool lifting = a.lifting != 0;
  const int ndir = span_split_ndir(lifting, a.Wa, a.Wb), len = span_split_row_len(lifting, a.Wa, a.Wb), nres = lifting ? 3 : 2;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(tid / ndir), dir = (int)(tid % ndir);
  if (i >= a.n) return;
  const bool wantJ = a.J && (a.flags & KTK_EVAL_JACOBIANS);
  if (!wantJ && dir != 0) return;
  const size_t dst = (size_t)((a.flags & KTK_EVAL_DEVICE_ORDER) ? i : a.perm[i]);
  const double ouv[2] = {a.obs_uv[2 * (size_t)i], a.obs_uv[2 * (size_t)i + 1]};
  const double obs_t0 = a.obs_t0[i];
  const int ridx = a.ref_idx[i];
  const int ka = span_window_base(a.sp.t0_r3, a.sp.dt_r3, obs_t0), kb = span_window_base(a.sp.t0_so3, a.sp.dt_so3, obs_t0);
  int st = kStatusRange, ira = -1, irb = -1;
  double r[3] = {nan(""), nan(""), nan("")};
  double* Jrow = wantJ ? a.J + dst * len : nullptr;
  if (ridx >= 0) {
    const double* rec = a.recs + (size_t)ridx * kRefSplitStride;
    ira = (int)rec[7]; irb = (int)rec[8];
    if (ira >= 0)
      st = span_split_column(lifting, a.sp, a.cam, a.vecs, a.quats, a.pairs, rec, ouv, obs_t0, a.ref_t0[i], lifting ? a.vt[i] : 0.0, ka, a.Wa, kb, a.Wb,
                             a.w[i], (a.flags & KTK_EVAL_ROBUST) ? a.huber[i] : 0.0, dir, r, Jrow);
  }
  if (st != 0) {
    atomicMin(a.err, st);
    r[0] = r[1] = r[2] = nan(""); ira = irb = -1;
    if (Jrow) { int stride; const int off = span_split_dir_offset(lifting, a.Wa, a.Wb, dir, stride); for (int rr = 0; rr < nres; ++rr) Jrow[off + rr * stride] = nan(""); }
  }
  if (dir == 0) {
    if (a.r) for (int rr = 0; rr < nres; ++rr) a.r[nres * dst + rr] = r[rr];
    if (a.idx[0]) a.idx[0][dst] = ira;
    if (a.idx[1]) a.idx[1][dst] = st == 0 ? ka : -1;
    if (a.idx[2]) a.idx[2][dst] = irb;
    if (a.idx[3]) a.idx[3][dst] = st == 0 ? kb : -1;
  }
}

// Sensor-block columns of NewtonRs / LiftingRs rows (newton_math.cuh "sensor-block columns"): one thread per (row, column 0..6) in forward mode.
// Js per row: [q_ct (nres x 4) | p_ct (nres x 3) | time offset (nres, zero)], nres = 2 / 3.  Cold path.
struct SpanSensorArgs {
  int traj; SplitConst spl; const double* vecs; const double* quats; const double* so3pairs; int Wa, Wb;      // split trajectory (traj == 1)
  SplineConst sp; CameraConst cam; const double* knots; const double* pairs; const double* rho;
  const double* obs_uv; const double* obs_t0; const double* ref_uv; const double* ref_t0; const int* lm; const double* w; const double* huber;
  const double* vt; const int* perm; int n, W, lifting; uint32_t flags; double* Js; int* err;
};
__global__ void k_span_sensor(const SpanSensorArgs a) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(tid / 7), c = (int)(tid % 7);
  if (i >= a.n) return;
  const int nres = a.lifting ? 3 : 2;
  const double ouv[2] = {a.obs_uv[2 * (size_t)i], a.obs_uv[2 * (size_t)i + 1]}, ruv[2] = {a.ref_uv[2 * (size_t)i], a.ref_uv[2 * (size_t)i + 1]};
  const double hub = (a.flags & KTK_EVAL_ROBUST) ? a.huber[i] : 0.0;
  double* dst = a.Js + (size_t)8 * nres * (size_t)((a.flags & KTK_EVAL_DEVICE_ORDER) ? i : a.perm[i]);
  int st;
  if (a.traj == 1) {
    const int ka = span_window_base(a.spl.t0_r3, a.spl.dt_r3, a.obs_t0[i]), kb = span_window_base(a.spl.t0_so3, a.spl.dt_so3, a.obs_t0[i]);
    st = span_split_sensor_column(a.lifting != 0, a.spl, a.cam, a.vecs, a.quats, a.so3pairs, ruv, a.ref_t0[i], a.rho[a.lm[i]], ouv, a.obs_t0[i],
                                  a.lifting ? a.vt[i] : 0.0, ka, a.Wa, kb, a.Wb, a.w[i], hub, c, dst);
  } else {
    const int kbase = newton_obs_window_base(a.sp, a.cam, a.obs_t0[i]);
    st = span_sensor_column(a.lifting != 0, a.sp, a.cam, a.knots, a.pairs, ruv, a.ref_t0[i], a.rho[a.lm[i]], ouv, a.obs_t0[i],
                            a.lifting ? a.vt[i] : 0.0, kbase, a.W, a.w[i], hub, c, dst);
  }
  if (st != 0) {
    atomicMin(a.err, st);
    for (int rr = 0; rr < nres; ++rr) { if (c < 4) dst[4 * rr + c] = nan(""); else dst[4 * nres + 3 * rr + (c - 4)] = nan(""); if (c == 0) dst[7 * nres + rr] = nan(""); }
  }
}

struct NewtonArgs {
  SplineConst sp; CameraConst cam;
  const double* knots; const double* pairs; const double* recs;
  const double* obs_uv; const double* obs_t0; const double* ref_t0; const int* ref_idx; const double* w; const double* huber;
  const int* perm;
  int n, W; uint32_t flags;
  double* r; double* J; int* i0r; int* i0o; int* err;
};
// Fast path of the Newton-RS rows (round 2).  The Newton iteration on the row time stops after its FIRST evaluation whenever the first step is
// below half a row (newton_rscamera_measurement.h:111-113) -- with half-pixel noise that is two rows out of three.  y_out is then the
// projection at the INITIAL row time, which does not depend on any parameter, so the Jet derivative the reference takes through the iteration
// is exactly the static row's Jacobian (bit for bit: tests/test_gpu_parity.py).  One thread per row runs the iteration once on values only;
// single-evaluation rows get the closed-form static row (written into the span layout), the others are flagged for the forward-mode kernel.
constexpr int kNewtonStage = 116;      // [Jref 56 | Jobs 56 | rho 2] + pad
constexpr int kNewtonRevSmemMax = 160 * 1024;      // k_newton_rs_rev stages d t_last / d theta (29 + 7 W doubles per row): W <= 86; wider spans take the dual-number kernels
__global__ void __launch_bounds__(32) k_newton_rs_fast(const NewtonArgs a, int* __restrict__ slow /* [0] = count, [1 + k] = row of the k-th slow row */,
                                                       double* __restrict__ slow_aux /* 6 per list slot: mode (0 forward mode, 2 two-evaluation row) | y(t_1) 2 | pi'(t_1) 2 */,
                                                       int allow_two) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31;
  double* row = smem + lane * kNewtonStage;
  const int i = blockIdx.x * 32 + lane;
  const bool wantJ = a.J && (a.flags & KTK_EVAL_JACOBIANS);
  const int row_len = 58 + 14 * a.W;
  int perm = -1, rel = 0;
  unsigned char slow_lane = 0;
  double aux[4] = {0.0, 0.0, 0.0, 0.0};
  int mode = -1;
  if (i < a.n) {
    const double ouv[2] = {a.obs_uv[2 * (size_t)i], a.obs_uv[2 * (size_t)i + 1]};
    const double obs_t0 = a.obs_t0[i], ref_t0 = a.ref_t0[i];
    const int ridx = a.ref_idx[i];
    const int kbase = newton_obs_window_base(a.sp, a.cam, obs_t0);
    if (ridx >= 0) {
      double r[2];
      int ir = -1;
      if (allow_two >= 2)
        mode = newton_rs_row_closed_any(a.sp, a.cam, a.knots, a.pairs, a.recs + (size_t)ridx * kRefStride, ouv, obs_t0, ref_t0, kbase, a.W, a.w[i],
                                        (a.flags & KTK_EVAL_ROBUST) ? a.huber[i] : 0.0, row, r, &ir, &rel, aux);
      else
        mode = newton_rs_row_closed(a.sp, a.cam, a.knots, a.pairs, a.recs + (size_t)ridx * kRefStride, ouv, obs_t0, ref_t0, kbase, a.W, a.w[i],
                                    (a.flags & KTK_EVAL_ROBUST) ? a.huber[i] : 0.0, allow_two != 0, row, r, &ir, &rel, aux);
      if (mode >= 0) {
        perm = (a.flags & KTK_EVAL_DEVICE_ORDER) ? i : a.perm[i];
        const size_t dst = (size_t)perm;
        if (a.r) { a.r[2 * dst] = r[0]; a.r[2 * dst + 1] = r[1]; }
        if (a.i0r) a.i0r[dst] = ir;
        if (a.i0o) a.i0o[dst] = kbase;
      }
    }
    slow_lane = mode != 0;      // forward mode (-1) and two-evaluation rows (2: their columns still get the d t_1 / d theta term) go on the list
  }
  {   // compact list of the rows the forward-mode kernel has to do (their order in the list does not matter: rows are independent)
    const unsigned m = __ballot_sync(0xffffffffu, slow_lane != 0);
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(slow, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    // ... and how many of them are forward-mode rows (slow[n + 1]): k_newton_rs leaves at once when there are none, instead of walking the list
    const unsigned mf = __ballot_sync(0xffffffffu, slow_lane != 0 && mode < 0);
    if (lane == 0 && mf) atomicAdd(slow + a.n + 1, __popc(mf));
    if (slow_lane) {
      const int slot = base + __popc(m & ((1u << lane) - 1u));
      slow[1 + slot] = i;
      double* ax = slow_aux + 6 * (size_t)slot;
      ax[0] = mode == 2 ? 2.0 : 0.0; ax[1] = aux[0]; ax[2] = aux[1]; ax[3] = aux[2]; ax[4] = aux[3]; ax[5] = 0.0;
    }
  }
  if (!wantJ) return;
  for (int rr = 0; rr < 32; ++rr) {
    const int d = __shfl_sync(0xffffffffu, perm, rr), sh = __shfl_sync(0xffffffffu, rel, rr);
    if (d < 0) continue;
    const double* src = smem + rr * kNewtonStage;
    double* dst = a.J + (size_t)d * row_len;
    const int lo = 56 + 14 * sh, tail0 = 56 + 14 * a.W;      // the four active blocks are one run of 56 doubles at `lo`
    for (int c = lane; c < row_len; c += 32) {
      double v;
      if (c < 56) v = src[c];
      else if (c >= tail0) v = src[112 + (c - tail0)];
      else v = (unsigned)(c - lo) < 56u ? src[56 + (c - lo)] : 0.0;
      dst[c] = v;
    }
  }
}

// Rows of the list whose iteration stopped after its second evaluation (mode 2): k_newton_rs_fast wrote the static row at t_1; every column gets
// + finish(pi'(t_1) d t_1 / d theta).  KTK_NEWTON_FAST=2 / 3 (kept as the cross-check of the closed form and for spans too wide for its staging):
// one WARP per listed row and 32 dual evaluations instead of 29 + 7 W (newton_math.cuh "newton_rs_first_step_lane"): lanes 0..27 the
// four observation knots active at the initial row time, lanes 28..31 the gradient of f/df with respect to the landmark X and rho, from which the 28
// reference-window columns and the rho column follow by the chain rule through the landmark record.  Lane L < 28 writes its observation column and
// reference column L, lane 28 the rho column: every column has one writer, no atomics.
__global__ void __launch_bounds__(128) k_newton_rs_two_w(const NewtonArgs a, const int* __restrict__ slow, const double* __restrict__ slow_aux) {
  if (!(a.J && (a.flags & KTK_EVAL_JACOBIANS))) return;
  const int row_len = 58 + 14 * a.W, lane = threadIdx.x & 31;
  const int nlist = slow[0], wstride = (int)((gridDim.x * blockDim.x) >> 5);
  for (int k = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); k < nlist; k += wstride) {
    const double* ax = slow_aux + 6 * (size_t)k;
    if (ax[0] != 2.0 || (ax[3] == 0.0 && ax[4] == 0.0)) continue;      // not a two-evaluation row / t_1 clamped: no correction (warp-uniform)
    const int i = slow[1 + k];
    const size_t dst = (size_t)((a.flags & KTK_EVAL_DEVICE_ORDER) ? i : a.perm[i]);
    const double ouv[2] = {a.obs_uv[2 * (size_t)i], a.obs_uv[2 * (size_t)i + 1]};
    const double obs_t0 = a.obs_t0[i], weight = a.w[i], huber_c = (a.flags & KTK_EVAL_ROBUST) ? a.huber[i] : 0.0;
    const int kbase = newton_obs_window_base(a.sp, a.cam, obs_t0);
    const double* rec = a.recs + (size_t)a.ref_idx[i] * kRefStride;
    int io = kbase;
    double dtd = 0.0;
    const int st = newton_rs_first_step_lane(a.sp, a.cam, a.knots, a.pairs, rec, ouv, obs_t0, a.ref_t0[i], kbase, a.W, lane, &io, dtd);
    double g[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) g[c] = __shfl_sync(0xffffffffu, dtd, 28 + c);
    double* Jr = a.J + dst * row_len;
    if (st != 0) {                                                        // the row time is outside the span for every lane alike
      if (lane == 0) atomicMin(a.err, st);
      for (int c = lane; c < row_len; c += 32) Jr[c] = nan("");
      continue;
    }
    double j[2];
    int stride;
    if (lane < 28) {
      newton_rs_two_step_finish(ouv, weight, huber_c, ax + 1, dtd, j);
      const int off = newton_dir_offset(28 + 7 * (io - kbase) + lane, a.W, stride);
      Jr[off] += j[0]; Jr[off + stride] += j[1];
    }
    if (lane < 29) {
      newton_rs_two_step_finish(ouv, weight, huber_c, ax + 1, newton_rs_ref_chain(rec, g, lane), j);
      const int off = newton_dir_offset(lane < 28 ? lane : 28 + 7 * a.W, a.W, stride);
      Jr[off] += j[0]; Jr[off + stride] += j[1];
    }
  }
}
// KTK_NEWTON_FAST >= 4: the listed rows (two or more evaluations) in CLOSED FORM, one thread per row: the iteration is run again on values with
// D = d t_last / d theta accumulated by one reverse sweep per evaluation (newton_math.cuh "NewtonRs rows in CLOSED FORM for any number of
// evaluations"), then the warp adds jfin (x) D to the static rows k_newton_rs_fast wrote, row after row with coalesced read-modify-writes.
__global__ void __launch_bounds__(32) k_newton_rs_rev(const NewtonArgs a, const int* __restrict__ slow, const double* __restrict__ slow_aux) {
  extern __shared__ __align__(16) double smem[];
  if (!(a.J && (a.flags & KTK_EVAL_JACOBIANS))) return;
  const int lane = threadIdx.x & 31, nD = 29 + 7 * a.W, dstride = nD | 1, row_len = 58 + 14 * a.W, tail0 = 56 + 14 * a.W;
  double* D = smem + lane * dstride;
  const int nlist = slow[0];
  for (int base = blockIdx.x * 32; base < nlist; base += gridDim.x * 32) {
    const int k = base + lane;
    long long dst = -1;
    double jf0 = 0.0, jf1 = 0.0;
    if (k < nlist) {
      const double* ax = slow_aux + 6 * (size_t)k;
      if (ax[0] == 2.0 && (ax[3] != 0.0 || ax[4] != 0.0)) {
        const int i = slow[1 + k];
        const double ouv[2] = {a.obs_uv[2 * (size_t)i], a.obs_uv[2 * (size_t)i + 1]};
        const double obs_t0 = a.obs_t0[i];
        const int kbase = newton_obs_window_base(a.sp, a.cam, obs_t0);
        NewtonIter it;
        const int st = newton_rs_iterate(a.sp, a.cam, a.knots, a.pairs, a.recs + (size_t)a.ref_idx[i] * kRefStride, ouv, obs_t0, a.ref_t0[i], kbase, a.W, it, D);
        dst = (a.flags & KTK_EVAL_DEVICE_ORDER) ? i : a.perm[i];
        if (st != 0) { atomicMin(a.err, st); jf0 = jf1 = nan(""); }
        else {
          double jfin[2];
          newton_rs_shift_column(it, ouv, a.w[i], (a.flags & KTK_EVAL_ROBUST) ? a.huber[i] : 0.0, jfin);
          jf0 = jfin[0]; jf1 = jfin[1];
        }
      }
    }
    __syncwarp();
    for (int rr = 0; rr < 32; ++rr) {
      const long long d = __shfl_sync(0xffffffffu, dst, rr);
      if (d < 0) continue;
      const double j0 = __shfl_sync(0xffffffffu, jf0, rr), j1 = __shfl_sync(0xffffffffu, jf1, rr);
      const double* Ds = smem + rr * dstride;
      double* Jr = a.J + (size_t)d * row_len;
      for (int c = lane; c < row_len; c += 32) {
        double v;
        if (c >= tail0) v = (c == tail0 ? j0 : j1) * Ds[nD - 1];
        else { const int blk = c / 14, w = c - 14 * blk; v = w < 7 ? j0 * Ds[7 * blk + w] : j1 * Ds[7 * blk + w - 7]; }
        Jr[c] += v;
      }
    }
    __syncwarp();
  }
}
__device__ __forceinline__ void newton_rs_item(const NewtonArgs& a, const int* __restrict__ slow, const double* __restrict__ slow_aux, long long tid) {
  const int ndir = 29 + 7 * a.W, row_len = 58 + 14 * a.W;
  int i = (int)(tid / ndir);
  const int dir = (int)(tid % ndir);
  if (i >= a.n) return;
  const double* ax = nullptr;
  if (slow) { if (i >= slow[0]) return; ax = slow_aux + 6 * (size_t)i; i = slow[1 + i]; }      // only the rows the fast path left (compact list)
  const bool wantJ = a.J && (a.flags & KTK_EVAL_JACOBIANS);
  if (!wantJ && dir != 0) return;
  const size_t dst = (size_t)((a.flags & KTK_EVAL_DEVICE_ORDER) ? i : a.perm[i]);
  const double ouv[2] = {a.obs_uv[2 * (size_t)i], a.obs_uv[2 * (size_t)i + 1]};
  const double obs_t0 = a.obs_t0[i];
  const int ridx = a.ref_idx[i];
  const int kbase = newton_obs_window_base(a.sp, a.cam, obs_t0);
  if (ax && ax[0] == 2.0) return;      // finished by k_newton_rs_fast + k_newton_rs_rev (or k_newton_rs_two_w)
  int st = kStatusRange;
  double r[2] = {nan(""), nan("")}, j[2] = {nan(""), nan("")};
  int ir = -1;
  if (ridx >= 0) {
    const double* rec = a.recs + (size_t)ridx * kRefStride;
    ir = (int)rec[7];
    if (ir >= 0) {
      NewtonRow o;
      st = newton_rs_direction(a.sp, a.cam, a.knots, a.pairs, rec, ouv, obs_t0, a.ref_t0[i], kbase, a.W, wantJ ? dir : -1, o);
      if (st == 0) newton_rs_finish(o, ouv, a.w[i], (a.flags & KTK_EVAL_ROBUST) ? a.huber[i] : 0.0, r, j);
    }
  }
  if (st != 0) { atomicMin(a.err, st); ir = -1; }
  if (dir == 0) {
    if (a.r) { a.r[2 * dst] = r[0]; a.r[2 * dst + 1] = r[1]; }
    if (a.i0r) a.i0r[dst] = ir;
    if (a.i0o) a.i0o[dst] = st == 0 ? kbase : -1;
  }
  if (wantJ) {
    int stride;
    const int off = newton_dir_offset(dir, a.W, stride);
    double* Jr = a.J + dst * row_len;
    Jr[off] = j[0]; Jr[off + stride] = j[1];
  }
}
__global__ void __launch_bounds__(128) k_newton_rs(const NewtonArgs a, const int* __restrict__ slow, const double* __restrict__ slow_aux) {
  if (slow && slow[a.n + 1] == 0) return;      // every listed row was finished in closed form
  const long long total = (long long)(slow ? slow[0] : a.n) * (29 + 7 * a.W);      // fixed grid striding over the device-side list (its length is only known on the device)
  for (long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x; tid < total; tid += (long long)gridDim.x * blockDim.x) newton_rs_item(a, slow, slow_aux, tid);
}

// LiftingRsCameraMeasurement rows in closed form (newton_math.cuh "LiftingRs rows in CLOSED FORM"): one thread per row like the static kernel,
// three residual rows through the same reverse sweep; vt[i] is the current value of row i's own parameter block.  A thread stages
// [Jref 84 | Jobs 84 | vt 3 | rho 3] in shared memory; the warp then writes the 32 packed rows [ref 4 x (3x7) | obs W x (3x7) | vt 3 | rho 3]
// cooperatively (8-byte stores, consecutive lanes on consecutive doubles), the active window at its place inside the W-knot span and zeros elsewhere.
constexpr int kLiftStage = 176;      // 174 staged doubles per row, padded to a 16-byte multiple
__global__ void __launch_bounds__(32) k_lifting_rs(const NewtonArgs a, const double* __restrict__ vt) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31;
  double* row = smem + lane * kLiftStage;
  const int i = blockIdx.x * 32 + lane;
  const bool wantJ = a.J && (a.flags & KTK_EVAL_JACOBIANS);
  const int row_len = 90 + 21 * a.W;
  int perm = -1, rel = 0;
  if (i < a.n) {
    perm = (a.flags & KTK_EVAL_DEVICE_ORDER) ? i : a.perm[i];
    const double ouv[2] = {a.obs_uv[2 * (size_t)i], a.obs_uv[2 * (size_t)i + 1]};
    const double obs_t0 = a.obs_t0[i];
    const int ridx = a.ref_idx[i];
    const int kbase = newton_obs_window_base(a.sp, a.cam, obs_t0);
    int st = kStatusRange, ir = -1, io = -1;
    double r[3] = {nan(""), nan(""), nan("")};
    if (ridx >= 0) {
      st = lifting_rs_row_analytic(a.sp, a.cam, a.knots, a.pairs, a.recs + (size_t)ridx * kRefStride, ouv, obs_t0, a.ref_t0[i], vt[i], a.w[i],
                                   (a.flags & KTK_EVAL_ROBUST) ? a.huber[i] : 0.0, r, row, row + 84, row + 168, row + 171, &ir, &io);
      if (st == 0 && (io < kbase || io + 4 > kbase + a.W)) st = kStatusRange;
    }
    if (st != 0) {
      atomicMin(a.err, st);
      r[0] = r[1] = r[2] = nan(""); ir = -1;
      KTK_COLD_LOOP for (int c = 0; c < 174; ++c) row[c] = nan("");
      io = kbase;
    }
    rel = io - kbase;
    const size_t dst = (size_t)perm;
    if (a.r) { a.r[3 * dst] = r[0]; a.r[3 * dst + 1] = r[1]; a.r[3 * dst + 2] = r[2]; }
    if (a.i0r) a.i0r[dst] = ir;
    if (a.i0o) a.i0o[dst] = st == 0 ? kbase : -1;
  }
  __syncwarp();
  if (!wantJ) return;
  for (int rr = 0; rr < 32; ++rr) {
    const int d = __shfl_sync(0xffffffffu, perm, rr), sh = __shfl_sync(0xffffffffu, rel, rr);
    if (d < 0) continue;
    const double* src = smem + rr * kLiftStage;
    double* dst = a.J + (size_t)d * row_len;
    const int lo = 84 + 21 * sh, tail0 = 84 + 21 * a.W;      // the four active blocks are one run of 84 doubles at `lo`; no division per element
    for (int c = lane; c < row_len; c += 32) {
      double v;
      if (c < 84) v = src[c];
      else if (c >= tail0) v = src[168 + (c - tail0)];
      else v = (unsigned)(c - lo) < 84u ? src[84 + (c - lo)] : 0.0;
      dst[c] = v;
    }
  }
}


// ---- host side ---------------------------------------------------------------------------------------------------
template <class T> struct DevBuf {
  T* p = nullptr; size_t n = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  int resize(size_t m) { if (m <= n && p) return KTK_OK; if (p) cudaFree(p); p = nullptr; n = 0; KTK_CUDA(cudaMalloc(&p, std::max<size_t>(m, 1) * sizeof(T))); n = m; return KTK_OK; }
  int upload(const std::vector<T>& v, cudaStream_t s) { int st = resize(v.size()); if (st) return st; if (!v.empty()) KTK_CUDA(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s)); return KTK_OK; }
};

struct Group {
  int kind = 0; int64_t n = 0;
  int ny = 3;                     // doubles of y per row: 3 (gyroscope / accelerometer / position), 4 (orientation: q)
  ktk_sensor sensor{}; ktk_camera cam{};
  int newton_W = 0;               // Newton-RS groups: knots of the widest observation span (0 = not computed for the current spline)
  int span_W[2] = {0, 0};         // ... on a split trajectory: R3 / SO3 spline
  // caller-order host copies (structure queries) and sorted device copies
  std::vector<double> t, y, w, obs_uv, obs_t0, ref_uv, ref_t0, huber;
  std::vector<int> lm, perm; int lm_max = -1, lm_min = 0;
  DevBuf<double> d_t, d_y, d_w, d_obs_uv, d_obs_t0, d_ref_t0, d_huber;
  DevBuf<int> d_perm, d_ref_idx, d_lm_caller, d_lm_sorted, d_io;
  DevBuf<double> d_uo;
  DevBuf<double> d_ref_uv_sorted;
  double bias[3] = {0.0, 0.0, 0.0};
  std::vector<double> vt; DevBuf<double> d_vt; bool vt_dirty = true;      // LiftingRs: current frame-normalised row times, caller order (ktk_set_group_vt)
  DevBuf<double> o_Js;
  DevBuf<double> o_amb; DevBuf<int> o_amb_i0, o_amb_i0b;      // span cameras, KTK_EVAL_LOCAL: ambient rows / window indices before k_span_localize
  DevBuf<int> d_slow;             // Newton-RS rows k_newton_rs_fast did not finish (more than one Newton evaluation): [count | row indices (n) | count of forward-mode rows]
  DevBuf<double> d_slow_aux;      // ... 6 doubles per list slot (k_newton_rs_fast)
  // landmark-reference records (static RS): one per distinct (landmark, segment origin of the reference evaluation)
  int64_t n_ref = 0;
  DevBuf<double> d_rr_uv, d_rr_t0, d_recs;
  DevBuf<int> d_rr_start, d_rr_n, d_rr_lm, d_rr_start_b, d_rr_n_b;
  DevBuf<int> o_i0c, o_i0d;
  // device-side outputs used by the host-buffer path
  DevBuf<double> o_r, o_J; DevBuf<int> o_i0, o_i0b;
  bool uploaded = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof;   // one event pair per profiled launch of this group's kernel
  ~Group() { for (auto& e : prof) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); } }
};

}  // namespace

namespace { struct GnState; void gn_state_free(GnState*); }
struct ktk_problem {
  GnState* gn = nullptr;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool have_spline = false;
  int traj = 0;                    // 0: UniformSE3SplineTrajectory, 1: SplitTrajectory (R3 + SO3)
  SplineConst sp{0.0, 1.0, 0, 0};
  SplitConst spl{0.0, 1.0, 0, 0.0, 1.0, 0};
  std::vector<Group*> groups;
  DevBuf<double> d_knots7, d_knots8, d_pairs, d_rho, d_vecs4, d_so3pairs;
  DevBuf<int> d_err;
  int* h_err = nullptr;   // pinned
  int64_t launches = 0;
  bool profiling = false;
  bool graphs_enabled = true;
  int cam_resident_tiles = 0;     // prefetch distance of k_static_rs (tiles), see ktk_problem_create
  int imu_resident_tiles = 0;     // ... of the IMU-row kernels
  int sm_count = 148;
  int newton_fast = 4;            // 4: every row in closed form whatever its number of evaluations: static row at t_last (k_newton_rs_fast) + jfin (x) d t_last / d theta
                                  //    by one reverse sweep per evaluation (k_newton_rs_rev); forward mode only where that fails (out-of-range rows);
                                  // 2, 3: rows that stop after one OR two evaluations in closed form; the latter + 32 dual evaluations per row (one warp per row,
                                  //    k_newton_rs_two_w);
                                  // 1: only one-evaluation rows; 0: every Newton-RS row through the forward-mode kernel (KTK_NEWTON_FAST, A/B and cross-check)
  // -1 (default): the IMU-like groups of an evaluation go out in ONE launch (k_short_batch) when the problem has no camera rows -- a chain of
  // one-wave kernels is launch-bound (C2 -2.6 %, C1: 3 kernels) -- and as one launch per group next to camera rows, where the fused launch measured
  // SLOWER under graph replay (H1 0.2648 -> 0.2571 ms, C4 0.2931 -> 0.2868 ms; profiles/r2x).  0 / 1: force; 2: + the landmark tables
  // (slower still).  KTK_FUSE_SHORT overrides (A/B).
  int fuse_short = -1;
  ShortBatch short_batch;
  cudaGraphExec_t graph_exec = nullptr;
  std::vector<uint64_t> graph_key;
  int64_t graph_launches = 0;
  ~ktk_problem() { gn_state_free(gn); if (graph_exec) cudaGraphExecDestroy(graph_exec); for (auto g : groups) delete g; if (h_err) cudaFreeHost(h_err); }
};

namespace {

void fill_sensor_consts(const ktk_sensor& s, ImuConst& c) {
  c.time_offset = s.time_offset; c.max_time_offset = s.max_time_offset; c.time_offset_locked = s.time_offset_locked;
  c.bias[0] = c.bias[1] = c.bias[2] = 0.0;
}

void fill_camera_consts(const ktk_camera& cm, CameraConst& c) {
  for (int i = 0; i < 9; ++i) c.K[i] = cm.K[i];
  // Eigen fixed-size 3x3 inverse = cofactors / determinant (pinhole_camera.h:63-67 inverts K per call)
  const double* a = cm.K; double k[9];
  k[0] = a[4] * a[8] - a[5] * a[7]; k[1] = a[2] * a[7] - a[1] * a[8]; k[2] = a[1] * a[5] - a[2] * a[4];
  k[3] = a[5] * a[6] - a[3] * a[8]; k[4] = a[0] * a[8] - a[2] * a[6]; k[5] = a[2] * a[3] - a[0] * a[5];
  k[6] = a[3] * a[7] - a[4] * a[6]; k[7] = a[1] * a[6] - a[0] * a[7]; k[8] = a[0] * a[4] - a[1] * a[3];
  const double det = a[0] * k[0] + a[1] * k[3] + a[2] * k[6];
  for (int i = 0; i < 9; ++i) c.Kinv[i] = k[i] / det;
  camera_set_pose(c, cm.base.q_ct, cm.base.p_ct);
  c.time_offset = cm.base.time_offset; c.max_time_offset = cm.base.max_time_offset; c.time_offset_locked = cm.base.time_offset_locked;
  c.readout = cm.readout; c.row_delta = cm.readout / (double)cm.rows;
  c.rows = cm.rows; c.model = cm.model; c.wc[0] = cm.wc[0]; c.wc[1] = cm.wc[1]; c.gamma = cm.gamma;
}

int check_sensor(const ktk_sensor* s) {
  if (!s) return fail(KTK_EINVAL, "sensor is NULL");
  if (!(s->max_time_offset >= 0.0)) return fail(KTK_EINVAL, "max_time_offset must be non-negative");
  return KTK_OK;
}

// Sort key = first active knot of the (first) spline evaluation; host arithmetic identical to the device's.
std::vector<int> sort_perm(const std::vector<int>& key) {
  std::vector<int> perm(key.size());
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return key[a] < key[b]; });
  return perm;
}
template <class T> std::vector<T> gather(const std::vector<T>& v, const std::vector<int>& perm, int width) {
  std::vector<T> o(v.size());
  for (size_t i = 0; i < perm.size(); ++i) for (int c = 0; c < width; ++c) o[i * width + c] = v[(size_t)perm[i] * width + c];
  return o;
}

int upload_group(ktk_problem* p, Group& g) {
  if (g.uploaded) return KTK_OK;
  if (!p->have_spline) return fail(KTK_EINVAL, "a trajectory (ktk_set_se3_spline / ktk_set_split_spline) must be set before evaluation");
  const bool split = p->traj == 1;
  // the spline whose first active knot orders the rows: SE3, or the SO3 part of a split trajectory
  const double kt0 = split ? p->spl.t0_so3 : p->sp.t0, kdt = split ? p->spl.dt_so3 : p->sp.dt;
  std::vector<int> key((size_t)g.n);
  if (is_camera(g.kind)) {
    const double row_delta = g.cam.readout / (double)g.cam.rows;
    for (int64_t i = 0; i < g.n; ++i) key[i] = knot_floor(g.obs_t0[i] + g.cam.base.time_offset + g.obs_uv[2 * i + 1] * row_delta, kt0, kdt);
  } else {
    for (int64_t i = 0; i < g.n; ++i) key[i] = knot_floor(g.t[i] + g.sensor.time_offset, kt0, kdt);
  }
  g.perm = sort_perm(key);
  g.d_lm_sorted.n = 0; g.d_lm_caller.n = 0;      // per-row landmark indices cached by the Gauss-Newton products: the row order just changed
  int st;
  cudaStream_t s = p->stream;
  if ((st = g.d_perm.upload(g.perm, s))) return st;
  if ((st = g.d_w.upload(gather(g.w, g.perm, 1), s))) return st;
  if (is_camera(g.kind)) {
    // Landmark-reference table.  The reference evaluation of a residual happens in the segment (of that residual's two
    // spans) that holds t_ref; its origin is what fixes (i0_ref, u_ref) bit-exactly.  Observations of one landmark
    // share it whenever the reference view is the earlier one; otherwise they get their own record.
    // (Split trajectory: one segment per spline, R3 in rr_start/rr_n and SO3 in rr_start_b/rr_n_b.)
    CameraConst cc; fill_camera_consts(g.cam, cc);
    std::unordered_map<uint64_t, int> index;
    std::vector<double> rr_uv, rr_t0; std::vector<int> rr_start, rr_n, rr_start_b, rr_n_b, rr_lm, ref_idx((size_t)g.n, -1);
    for (int64_t i = 0; i < g.n; ++i) {
      Segment s0{0, 0}, s1{0, 0}, sa{0, 0}, sb{0, 0};
      const double tr = static_rs_time(cc, g.ref_t0[i], g.ref_uv[2 * i + 1]);
      int ir; double ur;
      if (!split) {
        const int nseg = static_rs_segments(p->sp, cc, g.ref_t0[i], g.obs_t0[i], s0, s1);
        const int which = nseg == 0 ? -1 : locate_in_segments(nseg, s0, s1, tr, p->sp.t0, p->sp.dt, ir, ur);
        if (which < 0) continue;
        sa = which == 0 ? s0 : s1;
      } else {
        int nseg = static_rs_segments_split(p->spl, cc, g.ref_t0[i], g.obs_t0[i], p->spl.t0_r3, p->spl.dt_r3, s0, s1);
        int which = nseg == 0 ? -1 : locate_in_segments(nseg, s0, s1, tr, p->spl.t0_r3, p->spl.dt_r3, ir, ur);
        if (which < 0) continue;
        sa = which == 0 ? s0 : s1;
        nseg = static_rs_segments_split(p->spl, cc, g.ref_t0[i], g.obs_t0[i], p->spl.t0_so3, p->spl.dt_so3, s0, s1);
        which = nseg == 0 ? -1 : locate_in_segments(nseg, s0, s1, tr, p->spl.t0_so3, p->spl.dt_so3, ir, ur);
        if (which < 0) continue;
        sb = which == 0 ? s0 : s1;
      }
      if (g.lm[i] >= (1 << 24) || sa.start >= (1 << 20) || sb.start >= (1 << 20)) return fail(KTK_EUNSUPPORTED, "more than 2^24 landmarks or 2^20 knots");
      // two observations of a landmark always share (ref_uv, ref_t0): the landmark has ONE reference observation
      const uint64_t key = ((uint64_t)(uint32_t)g.lm[i] << 40) | ((uint64_t)(uint32_t)sa.start << 20) | (uint64_t)(uint32_t)sb.start;
      auto it = index.find(key);
      if (it == index.end()) {
        it = index.emplace(key, (int)rr_lm.size()).first;
        rr_uv.push_back(g.ref_uv[2 * i]); rr_uv.push_back(g.ref_uv[2 * i + 1]); rr_t0.push_back(g.ref_t0[i]);
        rr_start.push_back(sa.start); rr_n.push_back(sa.n); rr_start_b.push_back(sb.start); rr_n_b.push_back(sb.n); rr_lm.push_back(g.lm[i]);
      } else if (rr_uv[2 * it->second] != g.ref_uv[2 * i] || rr_uv[2 * it->second + 1] != g.ref_uv[2 * i + 1] || rr_t0[it->second] != g.ref_t0[i]) {
        return fail(KTK_EINVAL, "observations of one landmark disagree on its reference observation");
      }
      ref_idx[i] = it->second;
    }
    // order the records by their first knot (locality of the pair table), remap
    std::vector<int> rperm = sort_perm(split ? rr_start_b : rr_start), rinv(rperm.size());
    for (size_t k = 0; k < rperm.size(); ++k) rinv[rperm[k]] = (int)k;
    for (auto& v : ref_idx) if (v >= 0) v = rinv[v];
    g.n_ref = (int64_t)rr_lm.size();
    if ((st = g.d_rr_uv.upload(gather(rr_uv, rperm, 2), s))) return st;
    if ((st = g.d_rr_t0.upload(gather(rr_t0, rperm, 1), s))) return st;
    if ((st = g.d_rr_start.upload(gather(rr_start, rperm, 1), s))) return st;
    if ((st = g.d_rr_n.upload(gather(rr_n, rperm, 1), s))) return st;
    if ((st = g.d_rr_start_b.upload(gather(rr_start_b, rperm, 1), s))) return st;
    if ((st = g.d_rr_n_b.upload(gather(rr_n_b, rperm, 1), s))) return st;
    if ((st = g.d_rr_lm.upload(gather(rr_lm, rperm, 1), s))) return st;
    if ((st = g.d_recs.resize((size_t)g.n_ref * (split ? kRefSplitStride : kRefStride)))) return st;
    if ((st = g.d_ref_idx.upload(gather(ref_idx, g.perm, 1), s))) return st;
    if ((st = g.d_obs_uv.upload(gather(g.obs_uv, g.perm, 2), s))) return st;
    if ((st = g.d_obs_t0.upload(gather(g.obs_t0, g.perm, 1), s))) return st;
    if ((st = g.d_ref_t0.upload(gather(g.ref_t0, g.perm, 1), s))) return st;
    if ((st = g.d_huber.upload(gather(g.huber, g.perm, 1), s))) return st;
    if (!split) {      // observation-side lookup, independent of the evaluation point: once per row, here
      std::vector<int> io((size_t)g.n); std::vector<double> uo((size_t)g.n);
      for (int64_t i = 0; i < g.n; ++i) static_rs_row_locate_u(p->sp, cc, &g.obs_uv[2 * i], g.obs_t0[i], g.ref_t0[i], io[i], uo[i]);
      if ((st = g.d_io.upload(gather(io, g.perm, 1), s))) return st;
      if ((st = g.d_uo.upload(gather(uo, g.perm, 1), s))) return st;
    }
    if (!g.sensor.q_locked || !g.sensor.p_locked || !g.sensor.time_offset_locked) {     // inputs of the sensor-Jacobian kernel
      if ((st = g.d_ref_uv_sorted.upload(gather(g.ref_uv, g.perm, 2), s))) return st;
      if ((st = g.d_lm_sorted.upload(gather(g.lm, g.perm, 1), s))) return st;
    }
  } else {
    if ((st = g.d_t.upload(gather(g.t, g.perm, 1), s))) return st;
    if ((st = g.d_y.upload(gather(g.y, g.perm, g.ny), s))) return st;
  }
  KTK_CUDA(cudaStreamSynchronize(s));   // the gathered host vectors are temporaries
  g.uploaded = true;
  return KTK_OK;
}

// doubles per packed Jacobian row / per residual of a group (include/kontiki_b200.h "Layouts")
// Newton-RS groups: knots of the widest observation span {t0_obs - 1e-3, t0_obs + readout + 1e-3} on the current spline
// ... on a split trajectory: the widest span of the R3 (which = 0) / SO3 (1) spline
int span_window_split(const ktk_problem* p, const Group& g, int which) {
  int& cached = const_cast<Group&>(g).span_W[which];
  if (cached > 0) return cached;
  const double t0 = which == 0 ? p->spl.t0_r3 : p->spl.t0_so3, dt = which == 0 ? p->spl.dt_r3 : p->spl.dt_so3;
  int W = 4;
  for (int64_t i = 0; i < g.n; ++i) W = std::max(W, span_window_size(t0, dt, g.cam.readout, g.obs_t0[i]));
  cached = W;
  return W;
}
int newton_window(const ktk_problem* p, const Group& g) {
  if (g.newton_W > 0) return g.newton_W;
  CameraConst cc; fill_camera_consts(g.cam, cc);
  int W = 4;
  for (int64_t i = 0; i < g.n; ++i) W = std::max(W, newton_obs_window_size(p->sp, cc, g.obs_t0[i]));
  const_cast<Group&>(g).newton_W = W;
  return W;
}
int row_doubles(const ktk_problem* p, const Group& g, uint32_t flags = 0) {
  const bool local = (flags & KTK_EVAL_LOCAL) != 0;
  if (is_span_camera(g.kind) && p->traj == 1) {
    const int Wa = span_window_split(p, g, 0), Wb = span_window_split(p, g, 1), nres = g.kind == KTK_LIFTING_RS ? 3 : 2;
    return local ? nres * 3 * (8 + Wa + Wb) + (nres == 3 ? 6 : 2) : span_split_row_len(nres == 3, Wa, Wb);
  }
  if (g.kind == KTK_NEWTON_RS) return local ? 2 + 12 * (4 + newton_window(p, g)) : 58 + 14 * newton_window(p, g);
  if (g.kind == KTK_LIFTING_RS) return local ? 6 + 18 * (4 + newton_window(p, g)) : 90 + 21 * newton_window(p, g);
  if (g.kind == KTK_STATIC_RS) return local ? 98 : kCamRow;
  if (p->traj == 1 && g.kind == KTK_POSITION) return kPosSplitRow;
  if (g.kind == KTK_ORIENTATION) return p->traj == 1 ? (local ? 12 : kOriSplitRow) : (local ? 24 : 28);
  if (p->traj == 1) return g.kind == KTK_GYROSCOPE ? (local ? 36 : kGyroSplitRow) : (local ? 72 : kAccelSplitRow);
  return local ? 72 : kImuRow;
}
int res_doubles(const Group& g) { return g.kind == KTK_LIFTING_RS ? 3 : (is_camera(g.kind) ? 2 : (g.kind == KTK_ORIENTATION ? 1 : 3)); }

int add_imu(ktk_problem* p, int kind, const ktk_sensor* imu, int64_t n, const double* t, const double* y, const double* w) {
  if (!p) return fail(KTK_EINVAL, "problem is NULL");
  int st = check_sensor(imu); if (st) return st;
  if (n < 0 || (n > 0 && (!t || !y))) return fail(KTK_EINVAL, "bad measurement arrays");
  if (n > 0x7fffffff) return fail(KTK_EINVAL, "more than 2^31-1 measurements in one group");
  Group* g = new Group; g->kind = kind; g->n = n; g->sensor = *imu;
  g->ny = kind == KTK_ORIENTATION ? 4 : 3;
  g->t.assign(t, t + n); g->y.assign(y, y + g->ny * n);
  if (w) g->w.assign(w, w + n); else g->w.assign((size_t)n, 1.0);
  p->groups.push_back(g);
  if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
  p->graph_key.clear();
  return (int)p->groups.size() - 1;
}

}  // namespace

extern "C" {

const char* ktk_last_error(void) { return g_err.c_str(); }

int ktk_problem_create(int device, ktk_problem** out) {
  if (!out) return fail(KTK_EINVAL, "out is NULL");
  *out = nullptr;
  if (device < 0) { ktk_problem* hp = new ktk_problem; hp->device = -1; *out = hp; return KTK_OK; }   // structure queries only
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(KTK_ECUDA, std::string("no CUDA device (kontiki_b200 has no CPU path): ") + cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(KTK_EINVAL, "device index out of range");
  KTK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  KTK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(KTK_ECUDA, std::string("kontiki_b200 is built for sm_100a only; device is ") + prop.name);
  ktk_problem* p = new ktk_problem; p->device = device;
  int st = p->d_err.resize(1);
  if (st) { delete p; return st; }
  if (cudaHostAlloc(&p->h_err, sizeof(int), cudaHostAllocDefault) != cudaSuccess) { delete p; return fail(KTK_ECUDA, "cudaHostAlloc failed"); }
  // opt in to the shared-memory carve-out the row staging needs
  cudaFuncSetAttribute(k_imu<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kThreads * kImuRowStride * 8);
  cudaFuncSetAttribute(k_imu<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kThreads * kAccelRowStride * 8);
  cudaFuncSetAttribute(k_imu<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kThreads * kImuRowStride * 8);
  cudaFuncSetAttribute(k_imu<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kThreads * kImuRowStride * 8);
  cudaFuncSetAttribute(k_imu_split<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kThreads * kOriSplitStride * 8);
  cudaFuncSetAttribute(k_imu_split<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kThreads * kPosSplitStride * 8);
  cudaFuncSetAttribute(k_static_rs_local, cudaFuncAttributeMaxDynamicSharedMemorySize, (kCamThreads / 32) * kCamWarpSmem * 8);
  cudaFuncSetAttribute(k_landmark_ref, cudaFuncAttributeMaxDynamicSharedMemorySize, kThreads * kRefStride * 8);
  cudaFuncSetAttribute(k_lifting_rs, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * kLiftStage * 8);
  cudaFuncSetAttribute(k_static_rs, cudaFuncAttributeMaxDynamicSharedMemorySize, kCamThreads * kCamDevStride * 8);
  cudaFuncSetAttribute(k_short_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * kAccelRowStride * 8);
  if (const char* v = getenv("KTK_FUSE_SHORT")) p->fuse_short = atoi(v);
  if (const char* v = getenv("KTK_NEWTON_FAST")) p->newton_fast = atoi(v);
  cudaFuncSetAttribute(k_newton_rs_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * kNewtonStage * 8);
  cudaFuncSetAttribute(k_newton_rs_rev, cudaFuncAttributeMaxDynamicSharedMemorySize, kNewtonRevSmemMax);
  {   // tiles of k_static_rs resident on the chip = the distance of its input prefetch (7 warps x 148 SMs = 1036 on a B200)
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_static_rs, kCamThreads, kCamThreads * kCamDevStride * 8) == cudaSuccess)
      p->cam_resident_tiles = per_sm * (kCamThreads / 32) * prop.multiProcessorCount;
    if (const char* v = getenv("KTK_CAM_AHEAD")) p->cam_resident_tiles = atoi(v);      // A/B switch (0 = off)
    p->sm_count = prop.multiProcessorCount;
    p->imu_resident_tiles = 8 * prop.multiProcessorCount;                              // 8 one-warp CTAs per SM (register-limited)
    if (const char* v = getenv("KTK_IMU_AHEAD")) p->imu_resident_tiles = atoi(v);
  }
  cudaFuncSetAttribute(k_imu_split<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kThreads * kGyroSplitStride * 8);
  cudaFuncSetAttribute(k_imu_split<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kThreads * kAccelSplitStride * 8);
  cudaFuncSetAttribute(k_static_rs_split, cudaFuncAttributeMaxDynamicSharedMemorySize, (kCamThreads / 32) * kCamSplitWarpSmem * 8);
  cudaFuncSetAttribute(k_landmark_ref_split, cudaFuncAttributeMaxDynamicSharedMemorySize, kThreads * kRefSplitStride * 8);
  *out = p;
  return KTK_OK;
}

void ktk_problem_destroy(ktk_problem* p) { if (p) { if (p->device >= 0) cudaSetDevice(p->device); delete p; } }

static void drop_graph(ktk_problem* p) { if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; } p->graph_key.clear(); }
int ktk_set_stream(ktk_problem* p, void* s) { if (!p) return fail(KTK_EINVAL, "problem is NULL"); drop_graph(p); p->stream = (cudaStream_t)s; return KTK_OK; }
int ktk_set_graphs(ktk_problem* p, int32_t on) { if (!p) return fail(KTK_EINVAL, "problem is NULL"); drop_graph(p); p->graphs_enabled = on != 0; return KTK_OK; }

int ktk_set_se3_spline(ktk_problem* p, double dt, double t0, int32_t n_knots, int32_t compat) {
  if (!p) return fail(KTK_EINVAL, "problem is NULL");
  if (!(dt > 0.0)) return fail(KTK_EINVAL, "dt must be positive");
  if (n_knots < 4) return fail(KTK_ERANGE, "Spline had too few control points");   // spline_base.h:57-61
  p->sp.t0 = t0; p->sp.dt = dt; p->sp.n_knots = n_knots; p->sp.compat_zero_dB = compat;
  p->traj = 0; p->have_spline = true; drop_graph(p);
  for (auto g : p->groups) { g->uploaded = false; g->newton_W = 0; g->span_W[0] = g->span_W[1] = 0; g->vt_dirty = true; }   // the sort key depends on (t0, dt)
  return KTK_OK;
}

int ktk_set_split_spline(ktk_problem* p, double dt_r3, double t0_r3, int32_t n_r3, double dt_so3, double t0_so3, int32_t n_so3) {
  if (!p) return fail(KTK_EINVAL, "problem is NULL");
  if (!(dt_r3 > 0.0) || !(dt_so3 > 0.0)) return fail(KTK_EINVAL, "dt must be positive");
  if (n_r3 < 4 || n_so3 < 4) return fail(KTK_ERANGE, "Spline had too few control points");   // spline_base.h:57-61
  p->spl = SplitConst{t0_r3, dt_r3, n_r3, t0_so3, dt_so3, n_so3};
  p->traj = 1; p->have_spline = true; drop_graph(p);
  for (auto g : p->groups) { g->uploaded = false; g->newton_W = 0; g->span_W[0] = g->span_W[1] = 0; g->vt_dirty = true; }
  return KTK_OK;
}

int ktk_add_gyroscope(ktk_problem* p, const ktk_sensor* imu, int64_t n, const double* t, const double* y, const double* w) { return add_imu(p, KTK_GYROSCOPE, imu, n, t, y, w); }
int ktk_add_accelerometer(ktk_problem* p, const ktk_sensor* imu, int64_t n, const double* t, const double* y, const double* w) { return add_imu(p, KTK_ACCELEROMETER, imu, n, t, y, w); }
int ktk_add_position(ktk_problem* p, int64_t n, const double* t, const double* position, const double* w) {
  ktk_sensor none{};           // PositionMeasurement has no sensor: identity pose, zero locked time offset
  none.q_ct[3] = 1.0; none.max_time_offset = 0.0; none.q_locked = none.p_locked = none.time_offset_locked = 1;
  return add_imu(p, KTK_POSITION, &none, n, t, position, w);
}
int ktk_add_orientation(ktk_problem* p, int64_t n, const double* t, const double* q) {
  ktk_sensor none{};           // OrientationMeasurement has no sensor and no weight (orientation_measurement.h:20-31)
  none.q_ct[3] = 1.0; none.max_time_offset = 0.0; none.q_locked = none.p_locked = none.time_offset_locked = 1;
  return add_imu(p, KTK_ORIENTATION, &none, n, t, q, nullptr);
}

static int add_camera_group(ktk_problem* p, int kind, const ktk_camera* cam, int64_t n, const double* obs_uv, const double* obs_t0, const double* ref_uv,
                            const double* ref_t0, const int32_t* lm_idx, const double* w, const double* huber_c) {
  if (!p) return fail(KTK_EINVAL, "problem is NULL");
  if (!cam) return fail(KTK_EINVAL, "camera is NULL");
  int st = check_sensor(&cam->base); if (st) return st;
  if (cam->rows <= 0 || cam->cols <= 0) return fail(KTK_EINVAL, "camera rows/cols must be positive");
  if (cam->model != KTK_CAMERA_PINHOLE && cam->model != KTK_CAMERA_ATAN) return fail(KTK_EINVAL, "unknown camera model");
  if (cam->model == KTK_CAMERA_ATAN && !(cam->gamma != 0.0)) return fail(KTK_EINVAL, "AtanCamera needs gamma != 0");
  if (is_span_camera(kind) && !cam->base.time_offset_locked)      // (the relative pose may be unlocked: k_span_sensor)
    return fail(KTK_EUNSUPPORTED, "NewtonRs / LiftingRs camera measurements with an unlocked time offset are not built");
  if (n < 0 || (n > 0 && (!obs_uv || !obs_t0 || !ref_uv || !ref_t0 || !lm_idx))) return fail(KTK_EINVAL, "bad measurement arrays");
  if (n > 0x7fffffff) return fail(KTK_EINVAL, "more than 2^31-1 measurements in one group");
  Group* g = new Group; g->kind = kind; g->n = n; g->cam = *cam; g->sensor = cam->base;
  g->obs_uv.assign(obs_uv, obs_uv + 2 * n); g->obs_t0.assign(obs_t0, obs_t0 + n);
  g->ref_uv.assign(ref_uv, ref_uv + 2 * n); g->ref_t0.assign(ref_t0, ref_t0 + n);
  g->lm.assign(lm_idx, lm_idx + n);
  for (int l : g->lm) { g->lm_max = std::max(g->lm_max, l); g->lm_min = std::min(g->lm_min, l); }
  if (w) g->w.assign(w, w + n); else g->w.assign((size_t)n, 1.0);
  if (huber_c) g->huber.assign(huber_c, huber_c + n); else g->huber.assign((size_t)n, 5.0);
  if (kind == KTK_LIFTING_RS) { g->vt.resize((size_t)n); for (int64_t i = 0; i < n; ++i) g->vt[i] = obs_uv[2 * i + 1] / (double)cam->rows; }   // vt_orig, lifting_rscamera_measurement.h:68
  drop_graph(p);
  p->groups.push_back(g);
  return (int)p->groups.size() - 1;
}
int ktk_add_static_rs(ktk_problem* p, const ktk_camera* cam, int64_t n, const double* obs_uv, const double* obs_t0, const double* ref_uv,
                      const double* ref_t0, const int32_t* lm_idx, const double* w, const double* huber_c) {
  return add_camera_group(p, KTK_STATIC_RS, cam, n, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, w, huber_c);
}
int ktk_add_newton_rs(ktk_problem* p, const ktk_camera* cam, int64_t n, const double* obs_uv, const double* obs_t0, const double* ref_uv,
                      const double* ref_t0, const int32_t* lm_idx, const double* w, const double* huber_c) {
  return add_camera_group(p, KTK_NEWTON_RS, cam, n, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, w, huber_c);
}
int ktk_add_lifting_rs(ktk_problem* p, const ktk_camera* cam, int64_t n, const double* obs_uv, const double* obs_t0, const double* ref_uv,
                       const double* ref_t0, const int32_t* lm_idx, const double* w, const double* huber_c) {
  return add_camera_group(p, KTK_LIFTING_RS, cam, n, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, w, huber_c);
}
int ktk_set_group_vt(ktk_problem* p, int32_t group, const double* vt) {
  if (!p || group < 0 || group >= (int)p->groups.size() || !vt) return fail(KTK_EINVAL, "bad argument");
  Group& g = *p->groups[group];
  if (g.kind != KTK_LIFTING_RS) return fail(KTK_EINVAL, "only LiftingRs groups have a row-time parameter");
  g.vt.assign(vt, vt + g.n);
  g.vt_dirty = true;
  drop_graph(p);
  return KTK_OK;
}

int ktk_set_group_sensor(ktk_problem* p, int32_t group, const ktk_sensor* sensor) {
  if (!p || group < 0 || group >= (int)p->groups.size()) return fail(KTK_EINVAL, "bad group");
  int st = check_sensor(sensor); if (st) return st;
  Group& g = *p->groups[group];
  g.sensor = *sensor; g.cam.base = *sensor;
  g.uploaded = false; g.vt_dirty = true;            // the row order and the landmark-reference table depend on the time offset
  drop_graph(p);
  return KTK_OK;
}
int ktk_set_group_bias(ktk_problem* p, int32_t group, const double* bias) {
  if (!p || group < 0 || group >= (int)p->groups.size() || !bias) return fail(KTK_EINVAL, "bad argument");
  Group& g = *p->groups[group];
  if (is_camera(g.kind) || g.kind == KTK_POSITION || g.kind == KTK_ORIENTATION) return fail(KTK_EINVAL, "only IMU groups have a bias");
  for (int c = 0; c < 3; ++c) g.bias[c] = bias[c];
  drop_graph(p);
  return KTK_OK;
}

int32_t ktk_num_groups(const ktk_problem* p) { return p ? (int32_t)p->groups.size() : 0; }
int64_t ktk_group_size(const ktk_problem* p, int32_t g) { return (p && g >= 0 && g < (int)p->groups.size()) ? p->groups[g]->n : -1; }
int32_t ktk_group_kind(const ktk_problem* p, int32_t g) { return (p && g >= 0 && g < (int)p->groups.size()) ? p->groups[g]->kind : -1; }
int64_t ktk_launch_count(const ktk_problem* p) { return p ? p->launches : 0; }
int32_t ktk_group_row_size(const ktk_problem* p, int32_t g) { return (p && g >= 0 && g < (int)p->groups.size()) ? row_doubles(p, *p->groups[g]) : -1; }
int ktk_group_span_windows(const ktk_problem* p, int32_t g, int32_t* w_a, int32_t* w_b) {
  if (!p || g < 0 || g >= (int)p->groups.size() || !w_a || !w_b) return fail(KTK_EINVAL, "bad argument");
  const Group& grp = *p->groups[g];
  if (!is_span_camera(grp.kind)) return fail(KTK_EINVAL, "not a NewtonRs / LiftingRs group");
  if (!p->have_spline) return fail(KTK_EINVAL, "a trajectory must be set first");
  if (p->traj == 1) { *w_a = span_window_split(p, grp, 0); *w_b = span_window_split(p, grp, 1); }
  else { *w_a = newton_window(p, grp); *w_b = 0; }
  return KTK_OK;
}
int32_t ktk_group_row_size_local(const ktk_problem* p, int32_t g) { return (p && g >= 0 && g < (int)p->groups.size()) ? row_doubles(p, *p->groups[g], KTK_EVAL_LOCAL) : -1; }
int64_t ktk_num_knot_doubles(const ktk_problem* p) {
  if (!p || !p->have_spline) return 0;
  return p->traj == 1 ? (int64_t)3 * p->spl.n_r3 + (int64_t)4 * p->spl.n_so3 : (int64_t)7 * p->sp.n_knots;
}

// cold path: sensor-block Jacobians of one group (KTK_EVAL_SENSOR_JACOBIANS)
static int launch_sensor_jacobians(ktk_problem* p, Group& g, const ktk_group_out& o, uint32_t flags, const double* d_rho, const double* d_quats) {
  if (!(flags & KTK_EVAL_SENSOR_JACOBIANS) || !o.Js || g.n == 0) return KTK_OK;
  cudaStream_t s = p->stream;
  const int blocks = (int)((g.n + 127) / 128);
  if (g.kind == KTK_POSITION || g.kind == KTK_ORIENTATION) return KTK_OK;           // no sensor
  // a group whose sensor blocks are all locked has no sensor columns: the reference hands Ceres NULL Jacobians for constant blocks
  // (sensors.h:147-164).  The flag is per evaluation, not per group: an unlocked IMU next to a locked camera is the ordinary case.
  if (g.sensor.q_locked && g.sensor.p_locked && g.sensor.time_offset_locked) return KTK_OK;
  if (!is_camera(g.kind) && g.sensor.time_offset_locked) return KTK_OK;              // an IMU's relative pose is not applied (TODO.md:6): only the time offset has columns
  if (is_span_camera(g.kind)) {
    if (!g.sensor.time_offset_locked) return fail(KTK_EUNSUPPORTED, "an unlocked time offset under NewtonRs / LiftingRs camera measurements is not built (relative pose: yes)");
    if (g.d_ref_uv_sorted.n != (size_t)2 * g.n) return fail(KTK_EINVAL, "sensor Jacobians requested for a camera whose blocks are all locked");
    SpanSensorArgs a;
    a.traj = p->traj; a.spl = p->spl; a.vecs = p->d_vecs4.p; a.quats = d_quats; a.so3pairs = p->d_so3pairs.p;
    a.Wa = p->traj == 1 ? span_window_split(p, g, 0) : 0; a.Wb = p->traj == 1 ? span_window_split(p, g, 1) : 0;
    a.sp = p->sp; fill_camera_consts(g.cam, a.cam);
    a.knots = p->d_knots8.p; a.pairs = p->d_pairs.p; a.rho = d_rho;
    a.obs_uv = g.d_obs_uv.p; a.obs_t0 = g.d_obs_t0.p; a.ref_uv = g.d_ref_uv_sorted.p; a.ref_t0 = g.d_ref_t0.p; a.lm = g.d_lm_sorted.p; a.w = g.d_w.p;
    a.huber = g.d_huber.p; a.vt = g.d_vt.p; a.perm = g.d_perm.p; a.n = (int)g.n; a.W = p->traj == 1 ? 0 : newton_window(p, g); a.lifting = g.kind == KTK_LIFTING_RS;
    a.flags = flags; a.Js = o.Js; a.err = p->d_err.p;
    const long long threads = (long long)g.n * 7;
    k_span_sensor<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(a);
    p->launches += 1;
    return KTK_OK;
  }
  if (g.kind == KTK_STATIC_RS) {
    if (g.d_ref_uv_sorted.n != (size_t)2 * g.n) return fail(KTK_EINVAL, "sensor Jacobians requested for a camera whose blocks are all locked");
    CamSensorArgs a;
    a.traj = p->traj; a.spl = p->spl; a.vecs = p->d_vecs4.p; a.quats = d_quats; a.so3pairs = p->d_so3pairs.p;
    a.sp = p->sp; fill_camera_consts(g.cam, a.cam);
    a.knots = p->d_knots8.p; a.pairs = p->d_pairs.p; a.rho = d_rho;
    a.obs_uv = g.d_obs_uv.p; a.obs_t0 = g.d_obs_t0.p; a.ref_uv = g.d_ref_uv_sorted.p; a.ref_t0 = g.d_ref_t0.p; a.lm = g.d_lm_sorted.p; a.w = g.d_w.p;
    a.huber = g.d_huber.p; a.perm = g.d_perm.p; a.n = (int)g.n; a.flags = flags; a.Js = o.Js; a.err = p->d_err.p;
    k_static_rs_sensor<<<blocks, 128, 0, s>>>(a);
  } else {
    ImuSensorArgs a;
    a.traj = p->traj; a.which = g.kind == KTK_GYROSCOPE ? 0 : 1; a.sp = p->sp; a.spl = p->spl; fill_sensor_consts(g.sensor, a.imu);
    a.knots = p->d_knots8.p; a.pairs = p->d_pairs.p; a.vecs = p->d_vecs4.p; a.quats = d_quats; a.so3pairs = p->d_so3pairs.p;
    a.t = g.d_t.p; a.w = g.d_w.p; a.perm = g.d_perm.p; a.n = (int)g.n; a.Js = o.Js; a.err = p->d_err.p;
    a.device_order = (flags & KTK_EVAL_DEVICE_ORDER) ? 1 : 0;
    k_imu_sensor<<<blocks, 128, 0, s>>>(a);
  }
  p->launches += 1;
  return KTK_OK;
}

static int evaluate_device_se3(ktk_problem* p, const double* d_knots, const double* d_rho, uint32_t flags, const ktk_group_out* outs);

// ktk_evaluate_device for a split trajectory: d_knots = [R3 knots (3 n_r3) | SO3 knots (4 n_so3)], the parameter order of
// SplitEntity (split_trajectory.h:34-39, 117-123).
static int evaluate_device_split(ktk_problem* p, const double* d_knots, const double* d_rho, uint32_t flags, const ktk_group_out* outs) {
  cudaStream_t s = p->stream;
  const SplitConst& sp = p->spl;
  const double* d_quats = d_knots + (size_t)3 * sp.n_r3;      // 4-double records already
  k_pack_vecs<<<(sp.n_r3 * kVecStride + 255) / 256, 256, 0, s>>>(d_knots, sp.n_r3, p->d_vecs4.p, p->d_err.p);
  k_so3_pair_prepass<<<((sp.n_so3 - 1) * 9 + 127) / 128, 128, 0, s>>>(d_quats, sp.n_so3, p->d_so3pairs.p, p->d_err.p);
  p->launches += 2;
  for (size_t gi = 0; gi < p->groups.size(); ++gi) {
    Group& g = *p->groups[gi];
    if (g.n == 0) continue;
    const ktk_group_out& o = outs[gi];
    const int tpb = is_camera(g.kind) ? kCamThreads : kThreads;
    const int blocks = (int)((g.n + tpb - 1) / tpb);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (p->profiling) { KTK_CUDA(cudaEventCreate(&ev0)); KTK_CUDA(cudaEventCreate(&ev1)); g.prof.push_back({ev0, ev1}); KTK_CUDA(cudaEventRecord(ev0, s)); }
    if (is_camera(g.kind)) {
      RefSplitArgs ra;
      ra.sp = sp; fill_camera_consts(g.cam, ra.cam);
      ra.vecs = p->d_vecs4.p; ra.quats = d_quats; ra.pairs = p->d_so3pairs.p; ra.rho = d_rho;
      ra.ref_uv = g.d_rr_uv.p; ra.ref_t0 = g.d_rr_t0.p; ra.r3_start = g.d_rr_start.p; ra.r3_n = g.d_rr_n.p; ra.so3_start = g.d_rr_start_b.p; ra.so3_n = g.d_rr_n_b.p;
      ra.lm = g.d_rr_lm.p; ra.n = (int)g.n_ref; ra.recs = g.d_recs.p; ra.err = p->d_err.p;
      if (g.n_ref > 0) { k_landmark_ref_split<<<(int)((g.n_ref + kThreads - 1) / kThreads), kThreads, kThreads * kRefSplitStride * 8, s>>>(ra); p->launches += 1; }
      CamSplitArgs a;
      a.sp = sp; a.cam = ra.cam;
      a.vecs = p->d_vecs4.p; a.quats = d_quats; a.pairs = p->d_so3pairs.p; a.recs = g.d_recs.p;
      a.obs_uv = g.d_obs_uv.p; a.obs_t0 = g.d_obs_t0.p; a.ref_t0 = g.d_ref_t0.p; a.ref_idx = g.d_ref_idx.p; a.w = g.d_w.p; a.huber = g.d_huber.p;
      a.perm = g.d_perm.p; a.n = (int)g.n; a.flags = flags;
      a.r = o.r; a.J = o.J; a.idx[0] = o.i0; a.idx[1] = o.i0_b; a.idx[2] = o.i0_c; a.idx[3] = o.i0_d; a.err = p->d_err.p;
      if (is_span_camera(g.kind)) {
        SpanSplitArgs sa;
        sa.sp = sp; sa.cam = ra.cam; sa.vecs = a.vecs; sa.quats = a.quats; sa.pairs = a.pairs; sa.recs = a.recs;
        sa.obs_uv = a.obs_uv; sa.obs_t0 = a.obs_t0; sa.ref_t0 = a.ref_t0; sa.ref_idx = a.ref_idx; sa.w = a.w; sa.huber = a.huber; sa.vt = g.d_vt.p;
        sa.perm = a.perm; sa.n = a.n; sa.Wa = span_window_split(p, g, 0); sa.Wb = span_window_split(p, g, 1); sa.lifting = g.kind == KTK_LIFTING_RS;
        sa.flags = flags; sa.r = o.r; sa.J = o.J; for (int c = 0; c < 4; ++c) sa.idx[c] = a.idx[c]; sa.err = p->d_err.p;
        const bool localize = (flags & KTK_EVAL_LOCAL) && o.J && (flags & KTK_EVAL_JACOBIANS);
        if (localize) {      // ambient rows and the SO3 window indices into scratch; k_span_localize_split writes the caller's local rows
          sa.flags = flags & ~(uint32_t)KTK_EVAL_LOCAL; sa.J = g.o_amb.p;
          if (!sa.idx[2]) sa.idx[2] = g.o_amb_i0.p;
          if (!sa.idx[3]) sa.idx[3] = g.o_amb_i0b.p;
        }
        const long long threads = (long long)g.n * span_split_ndir(sa.lifting != 0, sa.Wa, sa.Wb);
        k_span_rs_split<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(sa);
        if (localize) {
          const int nres = sa.lifting ? 3 : 2;
          const long long th2 = (long long)g.n * (8 + sa.Wa + sa.Wb + 1);
          k_span_localize_split<<<(unsigned)((th2 + 127) / 128), 128, 0, s>>>(g.o_amb.p, sa.idx[2], sa.idx[3], d_quats, sp.n_so3, (int)g.n, sa.Wa, sa.Wb, nres,
                                                                            nres == 3 ? 6 : 2, o.J);
          p->launches += 1;
        }
      }
      else k_static_rs_split<<<blocks, kCamThreads, (kCamThreads / 32) * kCamSplitWarpSmem * 8, s>>>(a);
    } else {
      ImuSplitArgs a;
      a.sp = sp; fill_sensor_consts(g.sensor, a.imu);
      for (int c = 0; c < 3; ++c) a.imu.bias[c] = g.bias[c];
      a.vecs = p->d_vecs4.p; a.quats = d_quats; a.pairs = p->d_so3pairs.p;
      a.t = g.d_t.p; a.y = g.d_y.p; a.w = g.d_w.p; a.perm = g.d_perm.p; a.n = (int)g.n; a.flags = flags;
      a.r = o.r; a.J = o.J; a.i0_r3 = o.i0; a.i0_so3 = o.i0_c; a.err = p->d_err.p;
      if (g.kind == KTK_GYROSCOPE) k_imu_split<0><<<blocks, kThreads, kThreads * kGyroSplitStride * 8, s>>>(a);
      else if (g.kind == KTK_POSITION) k_imu_split<2><<<blocks, kThreads, kThreads * kPosSplitStride * 8, s>>>(a);
      else if (g.kind == KTK_ORIENTATION) k_imu_split<3><<<blocks, kThreads, kThreads * kOriSplitStride * 8, s>>>(a);
      else k_imu_split<1><<<blocks, kThreads, kThreads * kAccelSplitStride * 8, s>>>(a);
    }
    if (p->profiling) KTK_CUDA(cudaEventRecord(ev1, s));
    p->launches += 1;
    { const int sj = launch_sensor_jacobians(p, g, o, flags, d_rho, d_quats); if (sj) return sj; }
  }
  KTK_CUDA(cudaGetLastError());
  return KTK_OK;      // the status word is fetched by ktk_synchronize
}

int ktk_evaluate_device(ktk_problem* p, const double* d_knots, const double* d_rho, int64_t n_rho, uint32_t flags, const ktk_group_out* outs) {
  if (!p || !d_knots || !outs) return fail(KTK_EINVAL, "NULL argument");
  if (!p->have_spline) return fail(KTK_EINVAL, "a trajectory must be set before evaluation");
  if (p->device < 0) return fail(KTK_ECUDA, "this problem was created without a device (structure queries only); there is no CPU evaluation path");
  KTK_CUDA(cudaSetDevice(p->device));
  cudaStream_t s = p->stream;
  int st;
  for (size_t gi = 0; gi < p->groups.size(); ++gi)      // finished rows leave the SM through the TMA engine (cp.async.bulk): 16-byte aligned destinations
    if (outs[gi].J && (reinterpret_cast<uintptr_t>(outs[gi].J) & 15u)) return fail(KTK_EINVAL, "ktk_group_out.J must be 16-byte aligned (device pointer)");
  for (auto g : p->groups) {
    if ((st = upload_group(p, *g))) return st;
    if (g->kind == KTK_LIFTING_RS && g->n > 0 && (g->vt_dirty || g->d_vt.n != (size_t)g->n)) {      // the row times are part of the evaluation point (device order); outside any capture
      std::vector<double> sorted((size_t)g->n);
      for (int64_t k = 0; k < g->n; ++k) sorted[k] = g->vt[(size_t)g->perm[k]];
      if ((st = g->d_vt.upload(sorted, s))) return st;
      KTK_CUDA(cudaStreamSynchronize(s));
      g->vt_dirty = false;
    }
    if (g->kind == KTK_NEWTON_RS && g->n > 0 && ((st = g->d_slow.resize((size_t)g->n + 2)) || (st = g->d_slow_aux.resize((size_t)6 * g->n)))) return st;
    if (is_span_camera(g->kind) && (flags & KTK_EVAL_LOCAL) && (flags & KTK_EVAL_JACOBIANS) && g->n > 0) {      // scratch of k_span_localize (outside any capture)
      if ((st = g->o_amb.resize((size_t)g->n * row_doubles(p, *g)))) return st;
      if ((st = g->o_amb_i0.resize((size_t)g->n)) || (st = g->o_amb_i0b.resize((size_t)g->n))) return st;
    }
    if (is_camera(g->kind)) {
      if (!d_rho) return fail(KTK_EINVAL, "rho is NULL but the problem has camera measurements");
      if (g->n > 0 && (g->lm_min < 0 || g->lm_max >= n_rho)) return fail(KTK_EINVAL, "landmark index out of range of rho");
    }
  }
  // scratch buffers are sized outside any stream capture
  if (p->traj == 1) {
    if ((st = p->d_vecs4.resize((size_t)p->spl.n_r3 * kVecStride))) return st;
    if ((st = p->d_so3pairs.resize((size_t)p->spl.n_so3 * kSo3PairStride))) return st;
  } else {
    if ((st = p->d_knots8.resize((size_t)p->sp.n_knots * kKnotStride))) return st;
    if ((st = p->d_pairs.resize((size_t)p->sp.n_knots * kPairStride))) return st;
  }
  // CUDA graph: one evaluation is 5-8 small stream operations; with the same buffers as last time the whole sequence is
  // replayed as one graph launch (launch-bound for small problems: C1 is 12 us of kernels).  Event-timed runs are not captured.
  std::vector<uint64_t> key;
  const bool use_graph = p->graphs_enabled && !p->profiling && s != nullptr;
  if (use_graph) {
    key = {(uint64_t)flags, (uint64_t)d_knots, (uint64_t)d_rho, (uint64_t)n_rho, (uint64_t)p->traj};
    for (size_t gi = 0; gi < p->groups.size(); ++gi) {
      const ktk_group_out& o = outs[gi];
      for (const void* q : {(const void*)o.r, (const void*)o.J, (const void*)o.i0, (const void*)o.i0_b, (const void*)o.i0_c, (const void*)o.i0_d, (const void*)o.Js}) key.push_back((uint64_t)q);
      key.push_back((uint64_t)p->groups[gi]->n);
    }
    if (p->graph_exec && key == p->graph_key) {
      KTK_CUDA(cudaGraphLaunch(p->graph_exec, s));
      p->launches += p->graph_launches;
      return KTK_OK;
    }
    if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
    KTK_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  }
  const int64_t launches_before = p->launches;
  st = p->traj == 1 ? evaluate_device_split(p, d_knots, d_rho, flags, outs) : evaluate_device_se3(p, d_knots, d_rho, flags, outs);
  if (use_graph) {
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(s, &graph);
    if (st) { if (graph) cudaGraphDestroy(graph); return st; }
    if (ce != cudaSuccess) return fail(KTK_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    KTK_CUDA(cudaGraphInstantiate(&p->graph_exec, graph, 0));
    cudaGraphDestroy(graph);
    p->graph_key = key;
    p->graph_launches = p->launches - launches_before;
    KTK_CUDA(cudaGraphLaunch(p->graph_exec, s));
  }
  return st;
}

// One kernel after the other: running the short kernels (IMU-like rows, landmark records) as parallel branches of the graph was measured
// twice and lost both times (profiles/r1g_kernel_experiments.md): co-resident CTAs of different kernels share the instruction cache.
static int evaluate_device_se3(ktk_problem* p, const double* d_knots, const double* d_rho, uint32_t flags, const ktk_group_out* outs) {
  cudaStream_t s = p->stream;
  const int nk = p->sp.n_knots;
  k_pair_prepass<<<((nk - 1) * 15 + 127) / 128, 128, 0, s>>>(d_knots, nk, p->d_knots8.p, p->d_pairs.p, p->d_err.p);
  p->launches += 1;
  // the kernels of one group: "short" = IMU-like rows or the landmark records of a camera group (about one wave each on H1);
  // "rows" = the camera rows, which need their landmark records
  auto launch_short = [&](Group& g, const ktk_group_out& o, cudaStream_t ss) {
    if (is_camera(g.kind)) {
      RefArgs ra;
      ra.sp = p->sp; fill_camera_consts(g.cam, ra.cam);
      ra.knots = p->d_knots8.p; ra.pairs = p->d_pairs.p; ra.rho = d_rho;
      ra.ref_uv = g.d_rr_uv.p; ra.ref_t0 = g.d_rr_t0.p; ra.seg_start = g.d_rr_start.p; ra.seg_n = g.d_rr_n.p; ra.lm = g.d_rr_lm.p;
      ra.n = (int)g.n_ref; ra.recs = g.d_recs.p; ra.err = p->d_err.p;
      if (g.n_ref > 0) { k_landmark_ref<<<(int)((g.n_ref + kThreads - 1) / kThreads), kThreads, kThreads * kRefStride * 8, ss>>>(ra); p->launches += 1; }
      return;
    }
    const int blocks = (int)((g.n + kThreads - 1) / kThreads);
    ImuArgs a;
    a.sp = p->sp; fill_sensor_consts(g.sensor, a.imu);
    for (int c = 0; c < 3; ++c) a.imu.bias[c] = g.bias[c];
    a.knots = p->d_knots8.p; a.pairs = p->d_pairs.p;
    a.t = g.d_t.p; a.y = g.d_y.p; a.w = g.d_w.p; a.perm = g.d_perm.p; a.n = (int)g.n; a.flags = flags; a.ahead = p->imu_resident_tiles;
    a.r = o.r; a.J = o.J; a.i0 = o.i0; a.err = p->d_err.p;
    if (g.kind == KTK_GYROSCOPE) k_imu<0><<<blocks, kThreads, kThreads * kImuRowStride * 8, ss>>>(a);
    else if (g.kind == KTK_POSITION) k_imu<2><<<blocks, kThreads, kThreads * kImuRowStride * 8, ss>>>(a);
    else if (g.kind == KTK_ORIENTATION) k_imu<3><<<blocks, kThreads, kThreads * kImuRowStride * 8, ss>>>(a);
    else k_imu<1><<<blocks, kThreads, kThreads * kAccelRowStride * 8, ss>>>(a);
    p->launches += 1;
  };
  auto launch_rows = [&](Group& g, const ktk_group_out& o) {
    if (!is_camera(g.kind)) return;
    const int blocks = (int)((g.n + kCamThreads - 1) / kCamThreads);
    CamArgs a;
    a.sp = p->sp; fill_camera_consts(g.cam, a.cam);
    a.knots = p->d_knots8.p; a.pairs = p->d_pairs.p; a.recs = g.d_recs.p;
    a.obs_uv = g.d_obs_uv.p; a.obs_t0 = g.d_obs_t0.p; a.ref_t0 = g.d_ref_t0.p; a.ref_idx = g.d_ref_idx.p; a.w = g.d_w.p; a.huber = g.d_huber.p;
    a.perm = g.d_perm.p; a.io = g.d_io.p; a.uo = g.d_uo.p; a.n = (int)g.n; a.flags = flags; a.ahead = p->cam_resident_tiles;
    a.r = o.r; a.J = o.J; a.i0r = o.i0; a.i0o = o.i0_b; a.err = p->d_err.p;
    if (is_span_camera(g.kind)) {
      NewtonArgs na;
      na.sp = a.sp; na.cam = a.cam; na.knots = a.knots; na.pairs = a.pairs; na.recs = a.recs;
      na.obs_uv = a.obs_uv; na.obs_t0 = a.obs_t0; na.ref_t0 = a.ref_t0; na.ref_idx = a.ref_idx; na.w = a.w; na.huber = a.huber; na.perm = a.perm;
      na.n = a.n; na.W = newton_window(p, g); na.flags = flags; na.r = a.r; na.J = a.J; na.i0r = a.i0r; na.i0o = a.i0o; na.err = a.err;
      const bool localize = (flags & KTK_EVAL_LOCAL) && a.J && (flags & KTK_EVAL_JACOBIANS);
      if (localize) {      // ambient rows and the window indices into scratch; k_span_localize below writes the caller's local rows
        na.flags = flags & ~(uint32_t)KTK_EVAL_LOCAL; na.J = g.o_amb.p;
        if (!na.i0r) na.i0r = g.o_amb_i0.p;
        if (!na.i0o) na.i0o = g.o_amb_i0b.p;
      }
      if (g.kind == KTK_LIFTING_RS) {
        k_lifting_rs<<<(unsigned)((g.n + 31) / 32), 32, 32 * kLiftStage * 8, s>>>(na, g.d_vt.p);
      } else {
        const long long threads = (long long)g.n * (29 + 7 * na.W);
        const int rev_smem = 32 * ((29 + 7 * na.W) | 1) * 8;
        const int nf = (p->newton_fast >= 4 && rev_smem > kNewtonRevSmemMax) ? 3 : p->newton_fast;
        if (p->newton_fast) {
          cudaMemsetAsync(g.d_slow.p, 0, sizeof(int), s);
          cudaMemsetAsync(g.d_slow.p + g.n + 1, 0, sizeof(int), s);      // [0] listed rows, [n + 1] forward-mode rows among them
          k_newton_rs_fast<<<(unsigned)((g.n + 31) / 32), 32, 32 * kNewtonStage * 8, s>>>(na, g.d_slow.p, g.d_slow_aux.p, nf >= 4 ? 2 : nf >= 2 ? 1 : 0);
          p->launches += 1;
        }
        const unsigned grid = (unsigned)std::min<long long>((threads + 127) / 128, (long long)p->sm_count * 32);
        if (nf >= 4) {
          const unsigned gr = (unsigned)std::min<long long>((g.n + 31) / 32, (long long)p->sm_count * 16);
          k_newton_rs_rev<<<gr, 32, rev_smem, s>>>(na, g.d_slow.p, g.d_slow_aux.p); p->launches += 1;
        }
        else if (nf >= 2) { k_newton_rs_two_w<<<grid, 128, 0, s>>>(na, g.d_slow.p, g.d_slow_aux.p); p->launches += 1; }
        k_newton_rs<<<grid, 128, 0, s>>>(na, p->newton_fast ? g.d_slow.p : nullptr, g.d_slow_aux.p);
      }
      if (localize) {
        const int nres = g.kind == KTK_LIFTING_RS ? 3 : 2, tail = g.kind == KTK_LIFTING_RS ? 6 : 2;
        const long long threads = (long long)g.n * (4 + na.W + 1);
        k_span_localize<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(g.o_amb.p, na.i0r, na.i0o, p->d_knots8.p, p->sp.n_knots, (int)g.n, na.W, nres, tail, a.J);
        p->launches += 1;
      }
    }
    else if (!(flags & KTK_EVAL_LOCAL)) k_static_rs<<<blocks, kCamThreads, kCamThreads * kCamDevStride * 8, s>>>(a);
    else k_static_rs_local<<<blocks, kCamThreads, (kCamThreads / 32) * kCamWarpSmem * 8, s>>>(a);
    p->launches += 1;
  };
  // All short kernels in one launch (k_short_batch) when they fit its fixed-size argument block and no per-group event timing is wanted;
  // otherwise group by group as before.
  bool batched = false;
  int fuse_short = p->fuse_short;
  if (fuse_short < 0) {
    fuse_short = 1;
    for (auto g : p->groups) if (is_camera(g->kind) && g->n > 0) fuse_short = 0;
  }
  if (!p->profiling && fuse_short) {
    ShortBatch* b = &p->short_batch;
    b->n_imu = b->n_ref = 0;
    bool fits = true;
    int cta = 0;
    size_t smem = 0;
    for (int pass = 0; pass < 2 && fits; ++pass)
      for (size_t gi = 0; gi < p->groups.size() && fits; ++gi) {
        Group& g = *p->groups[gi];
        if (g.n == 0) continue;
        const ktk_group_out& o = outs[gi];
        if (pass == 0 && !is_camera(g.kind)) {
          if (b->n_imu == kShortMaxImu) { fits = false; break; }
          ImuArgs& a = b->imu[b->n_imu];
          a.sp = p->sp; fill_sensor_consts(g.sensor, a.imu);
          for (int c = 0; c < 3; ++c) a.imu.bias[c] = g.bias[c];
          a.knots = p->d_knots8.p; a.pairs = p->d_pairs.p;
          a.t = g.d_t.p; a.y = g.d_y.p; a.w = g.d_w.p; a.perm = g.d_perm.p; a.n = (int)g.n; a.flags = flags; a.ahead = p->imu_resident_tiles;
          a.r = o.r; a.J = o.J; a.i0 = o.i0; a.err = p->d_err.p;
          const int which = g.kind == KTK_GYROSCOPE ? 0 : (g.kind == KTK_POSITION ? 2 : (g.kind == KTK_ORIENTATION ? 3 : 1));
          b->which[b->n_imu] = which; b->first[b->n_imu] = cta; b->n_imu += 1;
          cta += (int)((g.n + 31) / 32);
          smem = std::max(smem, (size_t)32 * imu_stride(which) * 8);
        } else if (pass == 1 && fuse_short >= 2 && is_camera(g.kind) && g.n_ref > 0) {
          if (b->n_ref == kShortMaxRef) { fits = false; break; }
          RefArgs& ra = b->ref[b->n_ref];
          ra.sp = p->sp; fill_camera_consts(g.cam, ra.cam);
          ra.knots = p->d_knots8.p; ra.pairs = p->d_pairs.p; ra.rho = d_rho;
          ra.ref_uv = g.d_rr_uv.p; ra.ref_t0 = g.d_rr_t0.p; ra.seg_start = g.d_rr_start.p; ra.seg_n = g.d_rr_n.p; ra.lm = g.d_rr_lm.p;
          ra.n = (int)g.n_ref; ra.recs = g.d_recs.p; ra.err = p->d_err.p;
          b->first[b->n_imu + b->n_ref] = cta; b->n_ref += 1;
          cta += (int)((g.n_ref + 31) / 32);
          smem = std::max(smem, (size_t)32 * kRefStride * 8);
        }
      }
    if (fits && cta > 0 && b->n_imu + b->n_ref > 1) {
      b->first[b->n_imu + b->n_ref] = cta;
      k_short_batch<<<cta, 32, smem, s>>>(*b);
      p->launches += 1;
      batched = true;
    }
  }
  {
    for (size_t gi = 0; gi < p->groups.size(); ++gi) {
      Group& g = *p->groups[gi];
      if (g.n == 0) continue;
      const ktk_group_out& o = outs[gi];
      cudaEvent_t ev0 = nullptr, ev1 = nullptr;
      if (p->profiling) { KTK_CUDA(cudaEventCreate(&ev0)); KTK_CUDA(cudaEventCreate(&ev1)); g.prof.push_back({ev0, ev1}); KTK_CUDA(cudaEventRecord(ev0, s)); }
      if (!batched || (is_camera(g.kind) && fuse_short < 2)) launch_short(g, o, s);
      launch_rows(g, o);
      if (p->profiling) KTK_CUDA(cudaEventRecord(ev1, s));
      { const int sj = launch_sensor_jacobians(p, g, o, flags, d_rho, nullptr); if (sj) return sj; }
    }
  }
  KTK_CUDA(cudaGetLastError());
  return KTK_OK;      // the status word is fetched by ktk_synchronize: an evaluation is kernels only (C1: two graph nodes instead of four)
}

int ktk_synchronize(ktk_problem* p) {
  if (!p) return fail(KTK_EINVAL, "problem is NULL");
  if (p->device < 0) return fail(KTK_ECUDA, "problem has no device");
  KTK_CUDA(cudaSetDevice(p->device));
  KTK_CUDA(cudaMemcpyAsync(p->h_err, p->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
  KTK_CUDA(cudaStreamSynchronize(p->stream));
  const int e = *p->h_err;
  if (e == kStatusRange) return fail(KTK_ERANGE, "a measurement time is out of range for the trajectory (its output rows are NaN)");
  if (e == kStatusRuntime) return fail(KTK_ERUNTIME, "logq: Only implemented for unit quaternions (a SO3 knot pair is not unit norm)");
  if (e != 0) return fail(KTK_ERUNTIME, "device-side evaluation error");
  return KTK_OK;
}

int ktk_evaluate(ktk_problem* p, const double* knots, const double* rho, int64_t n_rho, uint32_t flags, const ktk_group_out* outs) {
  if (!p || !knots || !outs) return fail(KTK_EINVAL, "NULL argument");
  if (!p->have_spline) return fail(KTK_EINVAL, "a trajectory must be set before evaluation");
  if (p->device < 0) return fail(KTK_ECUDA, "this problem was created without a device (structure queries only); there is no CPU evaluation path");
  KTK_CUDA(cudaSetDevice(p->device));
  cudaStream_t s = p->stream;
  int st;
  const size_t nkd = (size_t)ktk_num_knot_doubles(p);
  if ((st = p->d_knots7.resize(nkd))) return st;
  KTK_CUDA(cudaMemcpyAsync(p->d_knots7.p, knots, nkd * sizeof(double), cudaMemcpyHostToDevice, s));
  if (rho && n_rho > 0) {
    if ((st = p->d_rho.resize((size_t)n_rho))) return st;
    KTK_CUDA(cudaMemcpyAsync(p->d_rho.p, rho, (size_t)n_rho * sizeof(double), cudaMemcpyHostToDevice, s));
  }
  std::vector<ktk_group_out> dev(p->groups.size());
  for (size_t gi = 0; gi < p->groups.size(); ++gi) {
    Group& g = *p->groups[gi];
    const ktk_group_out& o = outs[gi];
    const size_t n = (size_t)g.n;
    dev[gi] = ktk_group_out{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (o.r) { if ((st = g.o_r.resize(n * res_doubles(g)))) return st; dev[gi].r = g.o_r.p; }
    if (o.J && (flags & KTK_EVAL_JACOBIANS)) { if ((st = g.o_J.resize(n * row_doubles(p, g, flags)))) return st; dev[gi].J = g.o_J.p; }
    if (o.i0) { if ((st = g.o_i0.resize(n))) return st; dev[gi].i0 = g.o_i0.p; }
    if (o.i0_b) { if ((st = g.o_i0b.resize(n))) return st; dev[gi].i0_b = g.o_i0b.p; }
    if (o.i0_c) { if ((st = g.o_i0c.resize(n))) return st; dev[gi].i0_c = g.o_i0c.p; }
    if (o.i0_d) { if ((st = g.o_i0d.resize(n))) return st; dev[gi].i0_d = g.o_i0d.p; }
    if (o.Js && (flags & KTK_EVAL_SENSOR_JACOBIANS)) { if ((st = g.o_Js.resize(n * (g.kind == KTK_LIFTING_RS ? 24 : (is_camera(g.kind) ? 16 : 3))))) return st; dev[gi].Js = g.o_Js.p; }
  }
  if ((st = ktk_evaluate_device(p, p->d_knots7.p, (rho && n_rho > 0) ? p->d_rho.p : nullptr, n_rho, flags, dev.data()))) return st;
  for (size_t gi = 0; gi < p->groups.size(); ++gi) {
    Group& g = *p->groups[gi];
    const ktk_group_out& o = outs[gi];
    const size_t n = (size_t)g.n;
    const bool cam = is_camera(g.kind), split = p->traj == 1;
    if (dev[gi].r) KTK_CUDA(cudaMemcpyAsync(o.r, dev[gi].r, n * res_doubles(g) * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (dev[gi].J) KTK_CUDA(cudaMemcpyAsync(o.J, dev[gi].J, n * row_doubles(p, g, flags) * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (dev[gi].i0) KTK_CUDA(cudaMemcpyAsync(o.i0, dev[gi].i0, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (dev[gi].i0_b && cam) KTK_CUDA(cudaMemcpyAsync(o.i0_b, dev[gi].i0_b, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (dev[gi].i0_c && split) KTK_CUDA(cudaMemcpyAsync(o.i0_c, dev[gi].i0_c, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (dev[gi].i0_d && split && cam) KTK_CUDA(cudaMemcpyAsync(o.i0_d, dev[gi].i0_d, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (dev[gi].Js) KTK_CUDA(cudaMemcpyAsync(o.Js, dev[gi].Js, n * (g.kind == KTK_LIFTING_RS ? 24 : (cam ? 16 : 3)) * sizeof(double), cudaMemcpyDeviceToHost, s));
  }
  return ktk_synchronize(p);
}

static int traj_evaluate_impl(ktk_problem* p, const double* knots, int64_t n, const double* t, double* out, int32_t* status, int width);
int ktk_traj_evaluate(ktk_problem* p, const double* knots, int64_t n, const double* t, double* out, int32_t* status) {
  return traj_evaluate_impl(p, knots, n, t, out, status, 16);
}
int ktk_se3_evaluate_matrices(ktk_problem* p, const double* knots, int64_t n, const double* t, double* out, int32_t* status) {
  if (p && p->traj != 0) return fail(KTK_EINVAL, "not a UniformSE3SplineTrajectory");
  return traj_evaluate_impl(p, knots, n, t, out, status, 48);
}
static int traj_evaluate_impl(ktk_problem* p, const double* knots, int64_t n, const double* t, double* out, int32_t* status, int width) {
  if (!p || !knots || (n > 0 && (!t || !out || !status))) return fail(KTK_EINVAL, "NULL argument");
  if (!p->have_spline) return fail(KTK_EINVAL, "a trajectory must be set before evaluation");
  if (p->device < 0) return fail(KTK_ECUDA, "this problem was created without a device; there is no CPU evaluation path");
  if (n == 0) return KTK_OK;
  KTK_CUDA(cudaSetDevice(p->device));
  cudaStream_t s = p->stream;
  int st;
  const size_t nkd = (size_t)ktk_num_knot_doubles(p);
  DevBuf<double> d_t, d_out; DevBuf<int> d_st;
  if ((st = p->d_knots7.resize(nkd)) || (st = d_t.resize((size_t)n)) || (st = d_out.resize((size_t)n * width)) || (st = d_st.resize((size_t)n))) return st;
  KTK_CUDA(cudaMemcpyAsync(p->d_knots7.p, knots, nkd * sizeof(double), cudaMemcpyHostToDevice, s));
  KTK_CUDA(cudaMemcpyAsync(d_t.p, t, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
  KTK_CUDA(cudaMemsetAsync(p->d_err.p, 0, sizeof(int), s));
  const int blocks = (int)((n + 127) / 128);
  if (p->traj == 0) {
    const int nk = p->sp.n_knots;
    if ((st = p->d_knots8.resize((size_t)nk * kKnotStride)) || (st = p->d_pairs.resize((size_t)nk * kPairStride))) return st;
    k_pair_prepass<<<((nk - 1) * 15 + 127) / 128, 128, 0, s>>>(p->d_knots7.p, nk, p->d_knots8.p, p->d_pairs.p, nullptr);
    if (width == 16) k_traj_eval_se3<<<blocks, 128, 0, s>>>(p->sp, p->d_knots8.p, p->d_pairs.p, (int)n, d_t.p, d_out.p, d_st.p);
    else k_traj_eval_se3_matrices<<<blocks, 128, 0, s>>>(p->sp, p->d_knots8.p, p->d_pairs.p, (int)n, d_t.p, d_out.p, d_st.p);
  } else {
    const SplitConst& sp = p->spl;
    const double* d_quats = p->d_knots7.p + (size_t)3 * sp.n_r3;
    if ((st = p->d_vecs4.resize((size_t)sp.n_r3 * kVecStride)) || (st = p->d_so3pairs.resize((size_t)sp.n_so3 * kSo3PairStride))) return st;
    k_pack_vecs<<<(sp.n_r3 * kVecStride + 255) / 256, 256, 0, s>>>(p->d_knots7.p, sp.n_r3, p->d_vecs4.p, nullptr);
    k_so3_pair_prepass<<<((sp.n_so3 - 1) * 9 + 127) / 128, 128, 0, s>>>(d_quats, sp.n_so3, p->d_so3pairs.p, p->d_err.p);
    k_traj_eval_split<<<blocks, 128, 0, s>>>(sp, p->d_vecs4.p, d_quats, p->d_so3pairs.p, (int)n, d_t.p, d_out.p, d_st.p);
  }
  p->launches += p->traj == 0 ? 2 : 3;
  KTK_CUDA(cudaGetLastError());
  KTK_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)n * width * sizeof(double), cudaMemcpyDeviceToHost, s));
  KTK_CUDA(cudaMemcpyAsync(status, d_st.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s));
  KTK_CUDA(cudaMemcpyAsync(p->h_err, p->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  KTK_CUDA(cudaStreamSynchronize(s));
  if (*p->h_err == kStatusRuntime) return fail(KTK_ERUNTIME, "logq: Only implemented for unit quaternions (a SO3 knot pair is not unit norm)");
  for (int64_t i = 0; i < n; ++i) if (status[i] == kStatusRange) return fail(KTK_ERANGE, "t is out of range for the spline");
  return KTK_OK;
}

// ---- matrix-free Gauss-Newton products -------------------------------------------------------------------------------
static RowWindows row_windows(const ktk_problem* p, const Group& g) {
  RowWindows w{};
  const bool split = p->traj == 1;
  const int n_a = split ? p->spl.n_r3 : p->sp.n_knots, n_b = split ? p->spl.n_so3 : 0;
  const int colB = split ? 3 * n_a : 0;
  w.rho_col0 = split ? (long long)3 * n_a + (long long)4 * n_b : (long long)7 * n_a;
  w.rho_off_in_row = -1;
  w.nres = res_doubles(g);
  w.row_len = row_doubles(p, g);
  auto set = [&](int i, int off, int width, int col, int slot) { w.j_off[i] = off; w.width[i] = width; w.col_off[i] = col; w.slot[i] = slot; w.nk[i] = 4; };
  if (!split) {
    if (g.kind == KTK_NEWTON_RS) { w.nwin = 2; set(0, 0, 7, 0, 0); set(1, 56, 7, 0, 1); w.nk[1] = newton_window(p, g); w.rho_off_in_row = 56 + 14 * w.nk[1]; }
    else if (g.kind == KTK_STATIC_RS) { w.nwin = 2; set(0, 0, 7, 0, 0); set(1, 56, 7, 0, 1); w.rho_off_in_row = 112; }
    else { w.nwin = 1; set(0, 0, 7, 0, 0); }
  } else if (g.kind == KTK_STATIC_RS) {
    w.nwin = 4; set(0, 0, 3, 0, 0); set(1, 24, 4, colB, 2); set(2, 56, 3, 0, 1); set(3, 80, 4, colB, 3); w.rho_off_in_row = 112;
  } else if (g.kind == KTK_GYROSCOPE || g.kind == KTK_ORIENTATION) { w.nwin = 1; set(0, 0, 4, colB, 2); }
  else if (g.kind == KTK_POSITION) { w.nwin = 1; set(0, 0, 3, 0, 0); }
  else { w.nwin = 2; set(0, 0, 3, 0, 0); set(1, 36, 4, colB, 2); }
  return w;
}
int64_t ktk_num_parameters(const ktk_problem* p, int64_t n_rho) { return p ? ktk_num_knot_doubles(p) + n_rho : 0; }

// mode 0: u = J v ; mode 1: y += J^T u ; mode 2: y += diag(J^T J) ; mode 3: y_local += diag(P^T J^T J P)
static int apply_products(ktk_problem* p, int mode, uint32_t flags, const ktk_group_out* d_outs, const double* d_v, double* const* d_u, double* d_y,
                          const double* d_Pa = nullptr, const double* d_Pb = nullptr) {
  if (!p || !d_outs || (mode == 0 && (!d_v || !d_u)) || (mode == 1 && (!d_u || !d_y)) || (mode >= 2 && !d_y)) return fail(KTK_EINVAL, "NULL argument");
  if (p->device < 0) return fail(KTK_ECUDA, "problem has no device");
  KTK_CUDA(cudaSetDevice(p->device));
  cudaStream_t s = p->stream;
  for (size_t gi = 0; gi < p->groups.size(); ++gi) {
    Group& g = *p->groups[gi];
    if (g.n == 0) continue;
    const ktk_group_out& o = d_outs[gi];
    if (!o.J) return fail(KTK_EINVAL, "the group has no Jacobian rows in device memory");
    ApplyArgs a{};
    a.w = row_windows(p, g); a.n = (int)g.n; a.J = o.J;
    a.idx[0] = o.i0; a.idx[1] = o.i0_b; a.idx[2] = o.i0_c; a.idx[3] = o.i0_d;
    for (int w = 0; w < a.w.nwin; ++w) if (!a.idx[a.w.slot[w]]) return fail(KTK_EINVAL, "the group's index arrays are missing");
    if (mode == 3 && g.kind == KTK_NEWTON_RS) return fail(KTK_EUNSUPPORTED, "ktk_jtj_diagonal_local over NewtonRsCameraMeasurement rows is not built");
    if (g.kind == KTK_LIFTING_RS) return fail(KTK_EUNSUPPORTED, "the matrix-free products do not cover LiftingRs rows (their row-time blocks are per measurement)");
    if (is_span_camera(g.kind) && p->traj == 1) return fail(KTK_EUNSUPPORTED, "the matrix-free products do not cover NewtonRs rows on a split trajectory");
    if (is_camera(g.kind)) {        // landmark index of every row, in the order the rows were written
      if (flags & KTK_EVAL_DEVICE_ORDER) {
        if (g.perm.size() != (size_t)g.n) return fail(KTK_EINVAL, "device-order rows need an evaluation first");
        if (g.d_lm_sorted.n != (size_t)g.n) { int st = g.d_lm_sorted.upload(gather(g.lm, g.perm, 1), s); if (st) return st; KTK_CUDA(cudaStreamSynchronize(s)); }
        a.lm = g.d_lm_sorted.p;
      } else {
        if (g.d_lm_caller.n != (size_t)g.n) { int st = g.d_lm_caller.upload(g.lm, s); if (st) return st; KTK_CUDA(cudaStreamSynchronize(s)); }
        a.lm = g.d_lm_caller.p;
      }
    }
    a.v = d_v; a.u = d_u ? d_u[gi] : nullptr; a.y = d_y; a.mode = mode == 2 ? 1 : 0;
    if ((mode == 0 || mode == 1) && !a.u) return fail(KTK_EINVAL, "u is NULL for a non-empty group");
    const int blocks = (int)((g.n + 127) / 128);
    if (mode == 3) {
      const bool split = p->traj == 1;
      const long long n_a = split ? p->spl.n_r3 : p->sp.n_knots, n_b = split ? p->spl.n_so3 : 0;
      for (int w = 0; w < a.w.nwin; ++w) {
        const bool isB = split && a.w.width[w] == 4;
        a.P[w] = split ? (isB ? d_Pb : nullptr) : d_Pa;            // R3 knots are their own tangent space
        a.lwidth[w] = split ? 3 : 6;
        a.lcol_off[w] = isB ? 3 * n_a : 0;
        if ((!split || isB) && !a.P[w]) return fail(KTK_EINVAL, "tangent basis P is NULL");
      }
      a.lrho_col0 = split ? 3 * n_a + 3 * n_b : 6 * n_a;
      k_jtj_diag_local<<<blocks, 128, 0, s>>>(a);
    } else if (mode == 0) k_j_apply<<<blocks, 128, 0, s>>>(a); else k_jt_apply<<<blocks, 128, 0, s>>>(a);
    p->launches += 1;
  }
  KTK_CUDA(cudaGetLastError());
  return KTK_OK;
}
int ktk_j_apply(ktk_problem* p, uint32_t flags, const ktk_group_out* d_outs, const double* d_v, double* const* d_u) { return apply_products(p, 0, flags, d_outs, d_v, d_u, nullptr); }
int ktk_jt_apply(ktk_problem* p, uint32_t flags, const ktk_group_out* d_outs, double* const* d_u, double* d_y) { return apply_products(p, 1, flags, d_outs, nullptr, d_u, d_y); }
int ktk_jtj_diagonal(ktk_problem* p, uint32_t flags, const ktk_group_out* d_outs, double* d_y) { return apply_products(p, 2, flags, d_outs, nullptr, nullptr, d_y); }
int ktk_jtj_diagonal_local(ktk_problem* p, uint32_t flags, const ktk_group_out* d_outs, const double* d_Pa, const double* d_Pb, double* d_y) {
  return apply_products(p, 3, flags, d_outs, nullptr, nullptr, d_y, d_Pa, d_Pb);
}

int ktk_get_row_order(ktk_problem* p, int32_t group, int32_t* order) {
  if (!p || group < 0 || group >= (int)p->groups.size() || !order) return fail(KTK_EINVAL, "bad argument");
  if (!p->have_spline) return fail(KTK_EINVAL, "a trajectory must be set first");
  Group& g = *p->groups[group];
  if (g.perm.size() != (size_t)g.n || !g.uploaded) {          // the device order is fixed when the group is uploaded
    if (p->device < 0) return fail(KTK_ECUDA, "problem has no device");
    KTK_CUDA(cudaSetDevice(p->device));
    int st = upload_group(p, g); if (st) return st;
  }
  for (int64_t k = 0; k < g.n; ++k) order[k] = g.perm[k];
  return KTK_OK;
}

int ktk_set_profiling(ktk_problem* p, int32_t on) { if (!p) return fail(KTK_EINVAL, "problem is NULL"); p->profiling = on != 0; return KTK_OK; }

int ktk_read_profile(ktk_problem* p, int32_t group, double* total_ms, int64_t* launches) {
  if (!p || group < 0 || group >= (int)p->groups.size() || !total_ms || !launches) return fail(KTK_EINVAL, "bad argument");
  KTK_CUDA(cudaSetDevice(p->device));
  KTK_CUDA(cudaStreamSynchronize(p->stream));
  Group& g = *p->groups[group];
  double sum = 0.0;
  for (auto& e : g.prof) { float ms = 0.f; KTK_CUDA(cudaEventElapsedTime(&ms, e.first, e.second)); sum += ms; cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  *total_ms = sum; *launches = (int64_t)g.prof.size();
  g.prof.clear();
  return KTK_OK;
}

void* ktk_host_alloc(int64_t bytes) { void* ptr = nullptr; if (cudaHostAlloc(&ptr, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault) != cudaSuccess) { g_err = "cudaHostAlloc failed"; return nullptr; } return ptr; }
void ktk_host_free(void* ptr) { if (ptr) cudaFreeHost(ptr); }

// ---- structure (host only) ---------------------------------------------------------------------------------------
// *Measurement::AddToEstimator -> TrajectoryEstimator::AddTrajectoryForTimes -> SplineEntity::AddToProblem
// (gyroscope_measurement.h:82-92, static_rscamera_measurement.h:137-168, spline_base.h:361-404)
// which: 0 = the SE3 spline / the R3 part of a split trajectory, 1 = the SO3 part of a split trajectory
static int group_segments(const ktk_problem* p, const Group& g, int64_t i, int which, Segment& s0, Segment& s1) {
  const bool split = p->traj == 1;
  const double t0 = !split ? p->sp.t0 : (which == 0 ? p->spl.t0_r3 : p->spl.t0_so3);
  const double dt = !split ? p->sp.dt : (which == 0 ? p->spl.dt_r3 : p->spl.dt_so3);
  const double tmin = split ? split_min_time(p->spl) : p->sp.t0;
  const double tmax = split ? split_max_time(p->spl) : spline_max_time(p->sp);
  if (is_camera(g.kind)) {
    double t1, t2;
    if (g.ref_t0[i] <= g.obs_t0[i]) { t1 = g.ref_t0[i]; t2 = g.obs_t0[i]; } else { t1 = g.obs_t0[i]; t2 = g.ref_t0[i]; }
    if (!g.sensor.time_offset_locked) { t1 -= g.sensor.max_time_offset; t2 += g.sensor.max_time_offset; }
    const double margin = 1e-3;
    const double a1 = t1 - margin, b1 = t1 + g.cam.readout + margin, a2 = t2 - margin, b2 = t2 + g.cam.readout + margin;
    if (!(a1 >= tmin) || !(b1 < tmax) || !(a2 >= tmin) || !(b2 < tmax) || a1 > b1 || a2 > b2 || a2 < a1) return 0;
    return segments_two_spans(a1, b1, a2, b2, t0, dt, s0, s1);
  }
  double ta = g.t[i], tb = g.t[i];
  if (!g.sensor.time_offset_locked) { ta -= g.sensor.max_time_offset; tb += g.sensor.max_time_offset; }
  if (!(ta >= tmin) || !(tb < tmax)) return 0;
  segments_one_span(ta, tb, t0, dt, s0);
  return 1;
}

static int structure_of(const ktk_problem* p, int32_t group, int which, int32_t cap, int32_t* knot_ids, int32_t* n_ids) {
  if (!p || group < 0 || group >= (int)p->groups.size() || !knot_ids || !n_ids) return fail(KTK_EINVAL, "bad argument");
  if (!p->have_spline) return fail(KTK_EINVAL, "a trajectory must be set first");
  const Group& g = *p->groups[group];
  int worst = KTK_OK;
  for (int64_t i = 0; i < g.n; ++i) {
    Segment s0{0, 0}, s1{0, 0};
    const int nseg = group_segments(p, g, i, which, s0, s1);
    int32_t* ids = knot_ids + (size_t)i * cap;
    for (int c = 0; c < cap; ++c) ids[c] = -1;
    if (nseg == 0) { n_ids[i] = 0; worst = KTK_ERANGE; continue; }
    const int total = s0.n + (nseg == 2 ? s1.n : 0);
    if (total > cap) return fail(KTK_EINVAL, "cap too small for the knot blocks of a residual");
    int c = 0;
    for (int k = 0; k < s0.n; ++k) ids[c++] = s0.start + k;
    if (nseg == 2) for (int k = 0; k < s1.n; ++k) ids[c++] = s1.start + k;
    n_ids[i] = total;
  }
  if (worst == KTK_ERANGE) return fail(KTK_ERANGE, "Time span out of range for trajectory");
  return KTK_OK;
}

int ktk_get_structure(const ktk_problem* p, int32_t group, int32_t cap, int32_t* knot_ids, int32_t* n_ids) { return structure_of(p, group, 0, cap, knot_ids, n_ids); }
int ktk_get_structure_so3(const ktk_problem* p, int32_t group, int32_t cap, int32_t* knot_ids, int32_t* n_ids) {
  if (p && p->traj != 1) return fail(KTK_EINVAL, "not a split trajectory");
  return structure_of(p, group, 1, cap, knot_ids, n_ids);
}

int ktk_expand_static_rs(const ktk_problem* p, int32_t group, int32_t cap, const int32_t* knot_ids, const double* Jp, const int32_t* i0r,
                         const int32_t* i0o, double* out) {
  if (!p || group < 0 || group >= (int)p->groups.size() || !knot_ids || !Jp || !i0r || !i0o || !out) return fail(KTK_EINVAL, "bad argument");
  const Group& g = *p->groups[group];
  if (!is_camera(g.kind)) return fail(KTK_EINVAL, "not a camera group");
  if (p->traj != 0) return fail(KTK_EINVAL, "ktk_expand_static_rs is for the SE3 layout");
  const int W = is_span_camera(g.kind) ? newton_window(p, g) : 4, row_len = row_doubles(p, g);
  const int bs = 7 * res_doubles(g);      // doubles per knot block: 2 x 7, LiftingRs 3 x 7
  std::memset(out, 0, sizeof(double) * (size_t)g.n * cap * bs);
  for (int64_t i = 0; i < g.n; ++i) {
    const int32_t* ids = knot_ids + (size_t)i * cap;
    for (int w = 0; w < 2; ++w) {
      const int base = w == 0 ? i0r[i] : i0o[i];
      for (int k = 0; k < (w == 0 ? 4 : W); ++k) {
        const double* src = Jp + (size_t)i * row_len + w * 4 * bs + k * bs;
        int pos = -1;
        for (int c = 0; c < cap && ids[c] >= 0; ++c) if (ids[c] == base + k) { pos = c; break; }
        if (pos < 0) {       // a Newton-RS row carries the widest span of its group: blocks past this row's own span are zero
          bool zero = true;
          for (int c = 0; c < bs; ++c) zero = zero && src[c] == 0.0;
          if (zero && is_span_camera(g.kind)) continue;
          return fail(KTK_EINVAL, "active knot not in the structural block list");
        }
        double* dst = out + ((size_t)i * cap + pos) * bs;
        for (int c = 0; c < bs; ++c) dst[c] += src[c];
      }
    }
  }
  return KTK_OK;
}

}  // extern "C"

// =====================================================================================================================
// Gauss-Newton step on the device (gn_device.cuh): host orchestration and C ABI
// =====================================================================================================================
#include "gn_device.cuh"

namespace {

struct GnGroupState {
  int gi = 0, n = 0, nres = 0, row_len = 0, rho_off = -1;
  RowWindows rw{};
  const double* J = nullptr; const double* r = nullptr; const int* idx[4] = {nullptr, nullptr, nullptr, nullptr};
  DevBuf<int> order[4], fk[4], start[4];
  DevBuf<int> lm_order, lm_start; const int* lm_rows = nullptr;      // landmark of every row (device order)
  DevBuf<double> u, huber;
};
struct GnState {
  uint32_t flags = 0; int traj = 0;
  int n[2] = {0, 0}, width[2] = {0, 0}, lw[2] = {0, 0}, kind[2] = {0, 0}; int64_t n_rho = 0; double free_[2] = {1.0, 1.0};
  std::vector<GnGroupState*> g;
  struct View { double* p = nullptr; };      // a slice of one of the two exchange buffers below
  DevBuf<double> P[2], va[2], Minv[2], damp[2], x[2], r[2], p[2], b[2];
  DevBuf<double> cd, t, s, drho, partial, own;
  // what a sharded problem all-reduces lives in two contiguous buffers, so that each exchange is ONE collective:
  //   lin = [c | grho | blocks_a | blocks_b | z_a | z_b]  (once per linearisation),   qq = [q_a | q_b]  (reduced rhs; S p in every CG iteration)
  DevBuf<double> lin, qq; int64_t n_lin = 0, n_qq = 0;
  View c, grho, Bd[2], z[2], q[2];
  DevBuf<unsigned char> lm_locked; bool have_locked = false;
  DevBuf<GnScal> scal; GnScal* h_scal = nullptr;
  DevBuf<double> part; DevBuf<int> ticket;      // CG dot-product partials (4 x kGnCtas) and the last-CTA ticket
  GnWinList lists[2];
  const double* d_knots = nullptr; const double* d_rho = nullptr;
  ~GnState() { for (auto q : g) delete q; if (h_scal) cudaFreeHost(h_scal); }
};
void gn_state_free(GnState* s) { delete s; }

inline int gn_blocks(int64_t n, int per) { return (int)((n + per - 1) / per); }

GnVec gn_vec(GnState& S) {
  GnVec v{};
  for (int sp = 0; sp < 2; ++sp) {
    v.n[sp] = S.n[sp]; v.lw[sp] = S.lw[sp]; v.Minv[sp] = S.Minv[sp].p; v.damp[sp] = S.damp[sp].p; v.x[sp] = S.x[sp].p; v.r[sp] = S.r[sp].p;
    v.z[sp] = S.z[sp].p; v.p[sp] = S.p[sp].p; v.q[sp] = S.q[sp].p; v.b[sp] = S.b[sp].p;
  }
  return v;
}

// rows: u = J_k (P v) for every group, from local vectors vloc[2]
int gn_rows_apply(ktk_problem* p, GnState& S, double* const vloc[2]) {
  cudaStream_t s = p->stream;
  for (int sp = 0; sp < 2; ++sp)
    if (S.n[sp] > 0) k_gn_to_ambient<<<gn_blocks((int64_t)S.n[sp] * S.width[sp], 256), 256, 0, s>>>(S.P[sp].p, vloc[sp], S.n[sp], S.width[sp], S.lw[sp], S.free_[sp], S.va[sp].p);
  for (auto gp : S.g) {
    GnGroupState& G = *gp;
    GnRowsArgs a{};
    a.w = G.rw; a.n = G.n; a.J = G.J; for (int k = 0; k < 4; ++k) a.idx[k] = G.idx[k];
    a.va[0] = S.va[0].p; a.va[1] = S.va[1].p; a.u = G.u.p;
    k_gn_rows_apply<<<gn_blocks(G.n, 128), 128, 0, s>>>(a);
  }
  p->launches += 2 + (int64_t)S.g.size();
  return KTK_OK;
}
// t[l] = sum over all groups of J_rho^T x (x = u of the group, or r if from_r)
int gn_lm_sums(ktk_problem* p, GnState& S, bool from_r, bool squares, double* out) {
  cudaStream_t s = p->stream;
  if (S.n_rho == 0) return KTK_OK;
  bool first = true;
  for (auto gp : S.g) {
    GnGroupState& G = *gp;
    if (G.rho_off < 0) continue;
    k_gn_lm_reduce<<<gn_blocks(S.n_rho, 128), 128, 0, s>>>(G.J, G.row_len, G.rho_off, G.nres, squares ? nullptr : (from_r ? G.r : G.u.p), G.lm_order.p, G.lm_start.p,
                                                          (int)S.n_rho, first ? 0 : 1, out);
    first = false; p->launches += 1;
  }
  if (first) KTK_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)S.n_rho, s));
  return KTK_OK;
}
// u (or r -> u) minus J_rho s, every camera group
void gn_rows_fix(ktk_problem* p, GnState& S, bool from_r) {
  for (auto gp : S.g) {
    GnGroupState& G = *gp;
    if (G.rho_off < 0) { if (from_r) cudaMemcpyAsync(G.u.p, G.r, sizeof(double) * (size_t)G.n * G.nres, cudaMemcpyDeviceToDevice, p->stream); continue; }
    k_gn_rows_fix<<<gn_blocks(G.n, 256), 256, 0, p->stream>>>(G.J, G.row_len, G.rho_off, G.nres, G.lm_rows, S.s.p, G.n, from_r ? G.r : nullptr, G.u.p);
    p->launches += 1;
  }
}
void gn_gather(ktk_problem* p, GnState& S, double* const y[2]) {
  for (int sp = 0; sp < 2; ++sp)
    if (S.n[sp] > 0) {
      const int th = 32 * kGnWarpsPerKnot;
      if (S.width[sp] == 7) k_gn_gather<7><<<S.n[sp], th, 0, p->stream>>>(S.lists[sp], S.n[sp], S.lw[sp], S.P[sp].p, y[sp]);
      else if (S.width[sp] == 4) k_gn_gather<4><<<S.n[sp], th, 0, p->stream>>>(S.lists[sp], S.n[sp], S.lw[sp], S.P[sp].p, y[sp]);
      else k_gn_gather<3><<<S.n[sp], th, 0, p->stream>>>(S.lists[sp], S.n[sp], S.lw[sp], S.P[sp].p, y[sp]);
      p->launches += 1;
    }
}
void gn_set_x(GnState& S, bool use_r) {      // the gather transposes u (x = G.u) or the residuals
  for (int sp = 0; sp < 2; ++sp)
    for (int i = 0; i < S.lists[sp].n; ++i) {
      GnWinDev& w = S.lists[sp].w[i];
      for (auto gp : S.g) if (gp->J == w.J) w.x = use_r ? gp->r : gp->u.p;
    }
}

}  // namespace

extern "C" {

// Builds the static row lists (sorted by first knot per window, sorted by landmark) from the index arrays the LAST evaluation wrote into
// d_outs (rows in device order), and allocates the solver's vectors.  lm_locked: host array of n_rho bytes (1 = constant landmark) or NULL.
int ktk_gn_prepare(ktk_problem* p, uint32_t flags, const ktk_group_out* d_outs, int64_t n_rho, const uint8_t* lm_locked, int32_t lock_a, int32_t lock_b,
                   const double* const* huber_caller_order) {
  if (!p || !d_outs) return fail(KTK_EINVAL, "NULL argument");
  if (p->device < 0) return fail(KTK_ECUDA, "problem has no device");
  if (!(flags & KTK_EVAL_DEVICE_ORDER)) return fail(KTK_EINVAL, "the device Gauss-Newton step works on rows in device order (KTK_EVAL_DEVICE_ORDER)");
  if (flags & KTK_EVAL_LOCAL) return fail(KTK_EINVAL, "the device Gauss-Newton step takes ambient rows");
  KTK_CUDA(cudaSetDevice(p->device));
  cudaStream_t s = p->stream;
  KTK_CUDA(cudaStreamSynchronize(s));
  gn_state_free(p->gn); p->gn = nullptr;
  GnState* S = new GnState;
  p->gn = S;
  const bool split = p->traj == 1;
  S->flags = flags; S->traj = p->traj; S->n_rho = n_rho;
  S->n[0] = split ? p->spl.n_r3 : p->sp.n_knots; S->width[0] = split ? 3 : 7; S->lw[0] = split ? 3 : 6; S->kind[0] = split ? 1 : 0;
  S->n[1] = split ? p->spl.n_so3 : 0; S->width[1] = 4; S->lw[1] = 3; S->kind[1] = 2;
  S->free_[0] = lock_a ? 0.0 : 1.0; S->free_[1] = lock_b ? 0.0 : 1.0;
  S->lists[0].n = S->lists[1].n = 0;
  int st;
  for (size_t gi = 0; gi < p->groups.size(); ++gi) {
    Group& g = *p->groups[gi];
    if (g.n == 0) continue;
    if (is_span_camera(g.kind)) return fail(KTK_EUNSUPPORTED, "the device Gauss-Newton step does not cover NewtonRs / LiftingRs rows (host_cholesky solves them)");
    const ktk_group_out& o = d_outs[gi];
    if (!o.J || !o.r) return fail(KTK_EINVAL, "the group has no rows in device memory");
    GnGroupState* G = new GnGroupState;
    S->g.push_back(G);
    G->gi = (int)gi; G->n = (int)g.n; G->rw = row_windows(p, g); G->nres = G->rw.nres; G->row_len = G->rw.row_len; G->rho_off = G->rw.rho_off_in_row;
    G->J = o.J; G->r = o.r; G->idx[0] = o.i0; G->idx[1] = o.i0_b; G->idx[2] = o.i0_c; G->idx[3] = o.i0_d;
    if ((st = G->u.resize((size_t)G->n * G->nres))) return st;
    if (huber_caller_order && huber_caller_order[gi]) {
      std::vector<double> h((size_t)g.n);
      for (int64_t k = 0; k < g.n; ++k) h[k] = huber_caller_order[gi][g.perm[k]];
      if ((st = G->huber.upload(h, s))) return st;
    }
    // window lists
    std::vector<std::vector<int>> first(4);
    for (int w = 0; w < G->rw.nwin; ++w) {
      const int slot = G->rw.slot[w];
      if (!G->idx[slot]) return fail(KTK_EINVAL, "the group's index arrays are missing");
      if (first[slot].empty()) {
        first[slot].resize((size_t)g.n);
        KTK_CUDA(cudaMemcpy(first[slot].data(), G->idx[slot], sizeof(int) * (size_t)g.n, cudaMemcpyDeviceToHost));
      }
      const int sp = (split && G->rw.width[w] == 4) ? 1 : 0, nk = S->n[sp];
      std::vector<int> start((size_t)nk + 1, 0), order((size_t)g.n), fk((size_t)g.n);
      for (int64_t i = 0; i < g.n; ++i) {
        const int f = first[slot][i];
        if (f < 0 || f + 3 >= nk) return fail(KTK_ERANGE, "a row of the last evaluation has no valid knot window (evaluate successfully before ktk_gn_prepare)");
        start[(size_t)f + 1] += 1;
      }
      for (int k = 0; k < nk; ++k) start[(size_t)k + 1] += start[k];
      std::vector<int> cur(start.begin(), start.end() - 1);
      for (int64_t i = 0; i < g.n; ++i) { const int f = first[slot][i]; const int j = cur[f]++; order[j] = (int)i; fk[j] = f; }
      if ((st = G->order[w].upload(order, s)) || (st = G->fk[w].upload(fk, s)) || (st = G->start[w].upload(start, s))) return st;
      if (S->lists[sp].n >= kGnMaxWin) return fail(KTK_EUNSUPPORTED, "too many measurement groups for the device Gauss-Newton step");
      GnWinDev& d = S->lists[sp].w[S->lists[sp].n++];
      d.J = G->J; d.x = G->u.p; d.order = G->order[w].p; d.fk = G->fk[w].p; d.start = G->start[w].p;
      d.partner_first = nullptr; d.partner_j_off = 0; d.role = 0;
      d.j_off = G->rw.j_off[w]; d.width = G->rw.width[w]; d.nres = G->nres; d.row_len = G->row_len;
      if (g.kind == KTK_STATIC_RS) {      // windows come as (reference, observation) pairs on the same spline: w and w + nwin/2
        const int half = G->rw.nwin / 2, other = w < half ? w + half : w - half;
        d.role = w < half ? 1 : 2; d.partner_first = G->idx[G->rw.slot[other]]; d.partner_j_off = G->rw.j_off[other];
      }
    }
    if (is_camera(g.kind)) {
      if (g.lm_min < 0 || g.lm_max >= n_rho) return fail(KTK_EINVAL, "landmark index out of range of rho");
      if (g.d_lm_sorted.n != (size_t)g.n) { if ((st = g.d_lm_sorted.upload(gather(g.lm, g.perm, 1), s))) return st; }
      G->lm_rows = g.d_lm_sorted.p;
      std::vector<int> start((size_t)n_rho + 1, 0), order((size_t)g.n);
      for (int64_t i = 0; i < g.n; ++i) start[(size_t)g.lm[g.perm[i]] + 1] += 1;
      for (int64_t l = 0; l < n_rho; ++l) start[(size_t)l + 1] += start[l];
      std::vector<int> cur(start.begin(), start.end() - 1);
      for (int64_t i = 0; i < g.n; ++i) order[cur[g.lm[g.perm[i]]]++] = (int)i;
      if ((st = G->lm_order.upload(order, s)) || (st = G->lm_start.upload(start, s))) return st;
    }
  }
  for (int sp = 0; sp < 2; ++sp) {
    const size_t nk = (size_t)S->n[sp], lw = (size_t)S->lw[sp], wd = (size_t)S->width[sp];
    if ((st = S->P[sp].resize(nk * wd * lw)) || (st = S->va[sp].resize(nk * wd)) || (st = S->Minv[sp].resize(nk * lw * lw)) ||
        (st = S->damp[sp].resize(nk * lw)) || (st = S->x[sp].resize(nk * lw)) || (st = S->r[sp].resize(nk * lw)) ||
        (st = S->p[sp].resize(nk * lw)) || (st = S->b[sp].resize(nk * lw))) return st;
  }
  const size_t nr = (size_t)std::max<int64_t>(n_rho, 1);
  {
    const size_t kb[2] = {(size_t)S->n[0] * S->lw[0] * S->lw[0], (size_t)S->n[1] * S->lw[1] * S->lw[1]}, kv[2] = {(size_t)S->n[0] * S->lw[0], (size_t)S->n[1] * S->lw[1]};
    S->n_lin = (int64_t)(2 * nr + kb[0] + kb[1] + kv[0] + kv[1]); S->n_qq = (int64_t)(kv[0] + kv[1]);
    if ((st = S->lin.resize((size_t)S->n_lin + 1)) || (st = S->qq.resize((size_t)S->n_qq + 1))) return st;
    double* q = S->lin.p;
    S->c.p = q; q += nr; S->grho.p = q; q += nr; S->Bd[0].p = q; q += kb[0]; S->Bd[1].p = q; q += kb[1]; S->z[0].p = q; q += kv[0]; S->z[1].p = q;
    S->q[0].p = S->qq.p; S->q[1].p = S->qq.p + kv[0];
    KTK_CUDA(cudaMemsetAsync(S->lin.p, 0, sizeof(double) * (size_t)S->n_lin, s));
  }
  if ((st = S->cd.resize(nr)) || (st = S->t.resize(nr)) || (st = S->s.resize(nr)) || (st = S->drho.resize(nr)) || (st = S->own.resize(nr))) return st;
  int64_t maxrows = 1;
  for (auto gp : S->g) maxrows = std::max<int64_t>(maxrows, gp->n);
  if ((st = S->partial.resize((size_t)2 * gn_blocks(maxrows, 256) + 2))) return st;
  if ((st = S->scal.resize(1)) || (st = S->part.resize(4 * kGnCtas)) || (st = S->ticket.resize(1))) return st;
  KTK_CUDA(cudaMemsetAsync(S->ticket.p, 0, sizeof(int), s));
  if (cudaHostAlloc(&S->h_scal, sizeof(GnScal), cudaHostAllocDefault) != cudaSuccess) return fail(KTK_ECUDA, "cudaHostAlloc failed");
  if (lm_locked && n_rho > 0) {
    std::vector<unsigned char> l(lm_locked, lm_locked + n_rho);
    if ((st = S->lm_locked.upload(l, s))) return st;
    S->have_locked = true;
  }
  KTK_CUDA(cudaMemsetAsync(S->scal.p, 0, sizeof(GnScal), s));
  KTK_CUDA(cudaStreamSynchronize(s));
  KTK_CUDA(cudaGetLastError());
  return KTK_OK;
}

#define GN_STATE()                                                                                                            \
  if (!p || !p->gn) return fail(KTK_EINVAL, "ktk_gn_prepare has not been called");                                            \
  GnState& S = *p->gn;                                                                                                        \
  KTK_CUDA(cudaSetDevice(p->device));                                                                                         \
  cudaStream_t s = p->stream; (void)s

// cost 1/2 sum rho(s) of the rows in device memory -> scal.cost (this rank's rows)
int ktk_gn_cost(ktk_problem* p) {
  GN_STATE();
  bool first = true;
  for (auto gp : S.g) {
    GnGroupState& G = *gp;
    const int nb = gn_blocks(G.n, 256);
    const bool robust = (S.flags & KTK_EVAL_ROBUST) && G.huber.n == (size_t)G.n;
    k_gn_cost_rows<<<nb, 256, 0, s>>>(G.r, G.nres, robust ? G.huber.p : nullptr, G.n, S.partial.p);
    k_gn_sum_partials<<<1, 1024, 0, s>>>(S.partial.p, nb, 1, first ? 0 : 1, &S.scal.p->cost);
    first = false; p->launches += 2;
  }
  if (first) KTK_CUDA(cudaMemsetAsync(&S.scal.p->cost, 0, sizeof(double), s));
  return KTK_OK;
}

// Stage 0 (local to this rank's rows): tangent bases P, c_l = sum J_rho^2, g_rho = J_rho^T r, diagonal knot blocks B_kk.
// Between stages the caller all-reduces c, g_rho and the blocks when the rows are sharded over ranks (ktk_gn_buffer).
int ktk_gn_linearize_local(ktk_problem* p, const double* d_knots, const double* d_rho) {
  GN_STATE();
  if (!d_knots) return fail(KTK_EINVAL, "NULL argument");
  S.d_knots = d_knots; S.d_rho = d_rho;
  const double* kb = d_knots + (size_t)S.width[0] * S.n[0];
  k_gn_plus<<<gn_blocks(S.n[0], 128), 128, 0, s>>>(d_knots, S.n[0], S.kind[0], S.P[0].p);
  if (S.n[1] > 0) k_gn_plus<<<gn_blocks(S.n[1], 128), 128, 0, s>>>(kb, S.n[1], S.kind[1], S.P[1].p);
  int st;
  if ((st = gn_lm_sums(p, S, false, true, S.c.p)) || (st = gn_lm_sums(p, S, true, false, S.grho.p))) return st;
  if (S.n_rho > 0) KTK_CUDA(cudaMemcpyAsync(S.own.p, S.c.p, sizeof(double) * (size_t)S.n_rho, cudaMemcpyDeviceToDevice, s));      // > 0 where this rank holds rows of the landmark
  for (int sp = 0; sp < 2; ++sp)
    if (S.n[sp] > 0) {
      const int th = 32 * kGnWarpsPerKnot;
      if (S.width[sp] == 7) k_gn_blocks<7><<<S.n[sp], th, 0, s>>>(S.lists[sp], S.n[sp], S.lw[sp], S.P[sp].p, S.Bd[sp].p);
      else if (S.width[sp] == 4) k_gn_blocks<4><<<S.n[sp], th, 0, s>>>(S.lists[sp], S.n[sp], S.lw[sp], S.P[sp].p, S.Bd[sp].p);
      else k_gn_blocks<3><<<S.n[sp], th, 0, s>>>(S.lists[sp], S.n[sp], S.lw[sp], S.P[sp].p, S.Bd[sp].p);
      p->launches += 1;
    }
  p->launches += 2;
  KTK_CUDA(cudaGetLastError());
  return KTK_OK;
}
// g_k = P^T J_k^T r of this rank's rows (the knot part of the gradient, for the gradient-tolerance test) -> z buffers; g_rho is "grho".
int ktk_gn_gradient_local(ktk_problem* p) {
  GN_STATE();
  gn_set_x(S, true);
  double* y[2] = {S.z[0].p, S.z[1].p};
  gn_gather(p, S, y);
  gn_set_x(S, false);
  KTK_CUDA(cudaGetLastError());
  return KTK_OK;
}
__global__ void k_gn_damp_rho(const double* c, int n, double inv_radius, double* cd) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < n) cd[l] = c[l] > 0.0 ? c[l] + fmin(fmax(c[l], 1e-6), 1e32) * inv_radius : 0.0;
}
// Stage 1: LM damping of the landmark blocks, s = C^-1 g_rho, reduced gradient y = P^T J_k^T (r - J_rho s) of this rank's rows -> q buffers
// (all-reduce them before stage 2).
int ktk_gn_linearize_rhs(ktk_problem* p, double radius) {
  GN_STATE();
  if (S.n_rho > 0) {
    k_gn_damp_rho<<<gn_blocks(S.n_rho, 256), 256, 0, s>>>(S.c.p, (int)S.n_rho, 1.0 / radius, S.cd.p);
    k_gn_lm_scale<<<gn_blocks(S.n_rho, 256), 256, 0, s>>>(S.grho.p, S.cd.p, S.have_locked ? S.lm_locked.p : nullptr, (int)S.n_rho, S.s.p);
  }
  gn_rows_fix(p, S, true);
  gn_set_x(S, false);
  double* y[2] = {S.q[0].p, S.q[1].p};
  gn_gather(p, S, y);
  KTK_CUDA(cudaGetLastError());
  return KTK_OK;
}
// Stage 2: b = -y, block-Jacobi preconditioner (B_kk + D/radius)^-1, CG state x = 0.
int ktk_gn_pcg_begin(ktk_problem* p, double radius, double tol, int32_t max_iter) {
  GN_STATE();
  for (int sp = 0; sp < 2; ++sp)
    if (S.n[sp] > 0) {
      k_gn_negate<<<gn_blocks((int64_t)S.n[sp] * S.lw[sp], 256), 256, 0, s>>>(S.q[sp].p, S.n[sp] * S.lw[sp], S.free_[sp], S.b[sp].p);
      k_gn_invert_blocks<<<gn_blocks(S.n[sp], 64), 64, 0, s>>>(S.Bd[sp].p, S.n[sp], S.lw[sp], 1.0 / radius, S.free_[sp], S.Minv[sp].p, S.damp[sp].p);
    }
  k_gn_pcg_init<<<kGnCtas, kGnThreads, 0, s>>>(gn_vec(S), S.part.p);
  k_gn_pcg_init_scal<<<1, 1, 0, s>>>(S.part.p, S.scal.p, tol, max_iter);
  p->launches += 6;
  KTK_CUDA(cudaGetLastError());
  return KTK_OK;
}
// q = (S p) of this rank's rows, WITHOUT the damping term (ktk_gn_pcg_update adds D p after the all-reduce): 2 passes over the rows.
int ktk_gn_product(ktk_problem* p) {
  GN_STATE();
  double* pv[2] = {S.p[0].p, S.p[1].p};
  int st;
  if ((st = gn_rows_apply(p, S, pv))) return st;
  if (S.n_rho > 0) {
    if ((st = gn_lm_sums(p, S, false, false, S.t.p))) return st;
    k_gn_lm_scale<<<gn_blocks(S.n_rho, 256), 256, 0, s>>>(S.t.p, S.cd.p, S.have_locked ? S.lm_locked.p : nullptr, (int)S.n_rho, S.s.p);
    gn_rows_fix(p, S, false);
  }
  gn_set_x(S, false);
  double* y[2] = {S.q[0].p, S.q[1].p};
  gn_gather(p, S, y);
  KTK_CUDA(cudaGetLastError());
  return KTK_OK;
}
int ktk_gn_pcg_update(ktk_problem* p) {
  GN_STATE();
  const GnVec v = gn_vec(S);
  k_gn_pcg_a<<<kGnCtas, kGnThreads, 0, s>>>(v, S.scal.p, S.part.p);
  k_gn_pcg_b<<<kGnCtas, kGnThreads, 0, s>>>(v, S.scal.p, S.part.p, S.part.p + 2 * kGnCtas);
  k_gn_pcg_c<<<kGnCtas, kGnThreads, 0, s>>>(v, S.scal.p, S.part.p, S.part.p + 2 * kGnCtas, S.ticket.p);
  p->launches += 3;
  return KTK_OK;
}
// Synchronises and reports the CG state: iterations done, convergence flag, |r| / |b|.
int ktk_gn_pcg_status(ktk_problem* p, int32_t* iterations, int32_t* done, double* rel_residual) {
  GN_STATE();
  KTK_CUDA(cudaMemcpyAsync(S.h_scal, S.scal.p, sizeof(GnScal), cudaMemcpyDeviceToHost, s));
  KTK_CUDA(cudaStreamSynchronize(s));
  if (iterations) *iterations = S.h_scal->iter;
  if (done) *done = S.h_scal->done;
  if (rel_residual) *rel_residual = S.h_scal->bnorm2 > 0.0 ? sqrt(S.h_scal->rnorm2 / S.h_scal->bnorm2) : 0.0;
  return KTK_OK;
}
// After CG: delta_rho for the landmarks this rank holds (0 elsewhere: all-reduce the buffer "drho" when sharded), u = J delta of this rank's
// rows, and the partial sums of the model decrease -> scal.model_ur / model_uu (all-reduce the two when sharded), scal.step2 = |delta_k|^2.
int ktk_gn_finish_local(ktk_problem* p) {
  GN_STATE();
  double* xv[2] = {S.x[0].p, S.x[1].p};
  int st;
  if ((st = gn_rows_apply(p, S, xv))) return st;
  if (S.n_rho > 0) {
    if ((st = gn_lm_sums(p, S, false, false, S.t.p))) return st;
    k_gn_delta_rho<<<gn_blocks(S.n_rho, 256), 256, 0, s>>>(S.grho.p, S.t.p, S.cd.p, S.have_locked ? S.lm_locked.p : nullptr, (int)S.n_rho, S.drho.p);
  }
  k_gn_norm2<<<1, 1024, 0, s>>>(S.x[0].p, S.n[0] * S.lw[0], S.x[1].p, S.n[1] * S.lw[1], nullptr, 0, &S.scal.p->step2);
  KTK_CUDA(cudaGetLastError());
  return KTK_OK;
}
__global__ void k_gn_mask_own(const double* own, int n, double* d) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < n && !(own[l] > 0.0)) d[l] = 0.0;
}
int ktk_gn_finish_mask(ktk_problem* p) {      // zero the steps of landmarks held by other ranks (before the all-reduce that assembles delta_rho)
  GN_STATE();
  if (S.n_rho > 0) k_gn_mask_own<<<gn_blocks(S.n_rho, 256), 256, 0, s>>>(S.own.p, (int)S.n_rho, S.drho.p);
  return KTK_OK;
}
// With the complete delta_rho: model decrease sums over this rank's rows.
int ktk_gn_model_local(ktk_problem* p) {
  GN_STATE();
  bool first = true;
  for (auto gp : S.g) {
    GnGroupState& G = *gp;
    const int nb = gn_blocks(G.n, 256);
    k_gn_model_rows<<<nb, 256, 0, s>>>(G.J, G.row_len, G.rho_off, G.nres, G.lm_rows, S.drho.p, G.r, G.u.p, G.n, S.partial.p);
    k_gn_sum_partials<<<1, 1024, 0, s>>>(S.partial.p, nb, 2, first ? 0 : 1, &S.scal.p->model_ur);
    k_gn_sum_partials<<<1, 1024, 0, s>>>(S.partial.p + 1, nb, 2, first ? 0 : 1, &S.scal.p->model_uu);
    first = false; p->launches += 3;
  }
  KTK_CUDA(cudaGetLastError());
  return KTK_OK;
}
// Plus on the device: knots_out = Plus(knots_in, delta_k), rho_out = max(0, rho_in + delta_rho) (static_rscamera_measurement.h:180).
int ktk_gn_retract(ktk_problem* p, const double* d_knots_in, const double* d_rho_in, double* d_knots_out, double* d_rho_out) {
  GN_STATE();
  if (!d_knots_in || !d_knots_out) return fail(KTK_EINVAL, "NULL argument");
  const size_t offb = (size_t)S.width[0] * S.n[0];
  if (S.traj == 0) k_gn_retract_se3<<<gn_blocks(S.n[0], 128), 128, 0, s>>>(d_knots_in, S.x[0].p, S.n[0], d_knots_out);
  else {
    k_gn_retract_add<<<gn_blocks((int64_t)3 * S.n[0], 256), 256, 0, s>>>(d_knots_in, S.x[0].p, 3 * S.n[0], 0.0, 0, d_knots_out);
    k_gn_retract_so3<<<gn_blocks(S.n[1], 128), 128, 0, s>>>(d_knots_in + offb, S.x[1].p, S.n[1], d_knots_out + offb);
  }
  if (S.n_rho > 0 && d_rho_in && d_rho_out) k_gn_retract_add<<<gn_blocks(S.n_rho, 256), 256, 0, s>>>(d_rho_in, S.drho.p, (int)S.n_rho, 0.0, 1, d_rho_out);
  p->launches += 3;
  KTK_CUDA(cudaGetLastError());
  return KTK_OK;
}
// Device buffers of the solver, for the all-reduces of a sharded problem and for tests.  Returns the element count (doubles), 0 if unknown.
int64_t ktk_gn_buffer(ktk_problem* p, const char* name, double** ptr) {
  if (!p || !p->gn || !name || !ptr) return 0;
  GnState& S = *p->gn;
  const std::string n(name);
  auto kn = [&](int sp, int per) { return (int64_t)S.n[sp] * per; };
  if (n == "lin") { *ptr = S.lin.p; return S.n_lin; }
  if (n == "qq") { *ptr = S.qq.p; return S.n_qq; }
  if (n == "c") { *ptr = S.c.p; return S.n_rho; }
  if (n == "grho") { *ptr = S.grho.p; return S.n_rho; }
  if (n == "drho") { *ptr = S.drho.p; return S.n_rho; }
  if (n == "blocks_a") { *ptr = S.Bd[0].p; return kn(0, S.lw[0] * S.lw[0]); }
  if (n == "blocks_b") { *ptr = S.Bd[1].p; return kn(1, S.lw[1] * S.lw[1]); }
  if (n == "q_a") { *ptr = S.q[0].p; return kn(0, S.lw[0]); }
  if (n == "q_b") { *ptr = S.q[1].p; return kn(1, S.lw[1]); }
  if (n == "x_a") { *ptr = S.x[0].p; return kn(0, S.lw[0]); }
  if (n == "x_b") { *ptr = S.x[1].p; return kn(1, S.lw[1]); }
  if (n == "p_a") { *ptr = S.p[0].p; return kn(0, S.lw[0]); }
  if (n == "p_b") { *ptr = S.p[1].p; return kn(1, S.lw[1]); }
  if (n == "z_a") { *ptr = S.z[0].p; return kn(0, S.lw[0]); }
  if (n == "z_b") { *ptr = S.z[1].p; return kn(1, S.lw[1]); }
  if (n == "b_a") { *ptr = S.b[0].p; return kn(0, S.lw[0]); }
  if (n == "b_b") { *ptr = S.b[1].p; return kn(1, S.lw[1]); }
  if (n == "scal") { *ptr = reinterpret_cast<double*>(S.scal.p); return (int64_t)(sizeof(GnScal) / sizeof(double)); }
  return 0;
}

}  // extern "C"
#ifdef KTK_PHASE_TIMING
extern "C" int ktk_debug_read_phases(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out, g_phase, sizeof(unsigned long long) * 8) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; cudaMemcpyToSymbol(g_phase, z, sizeof(z)); }
  return 0;
}
#endif
