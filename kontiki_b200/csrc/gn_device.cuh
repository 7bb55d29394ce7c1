// kontiki_b200 -- Gauss-Newton / Levenberg-Marquardt step on the device, included by ktk.cu.
//
// The reference hands the normal equations to Ceres (SPARSE_SCHUR, cpplib/include/kontiki/trajectory_estimator.h:38-64): Ceres eliminates
// the landmark blocks (here 1x1: the inverse depth rho, static_rscamera_measurement.h:178-184) and factorises the reduced knot system.
// Here the rows an evaluation left in device memory (KTK_EVAL_DEVICE_ORDER) are never copied or assembled:
//   * rho is eliminated on the device: c_l = sum J_rho^2 + damping, the Schur complement S = B - E C^-1 E^T is APPLIED (implicit Schur):
//       S v = J_k^T ( u - J_rho C^-1 J_rho^T u ) + D v,   u = J_k P v        (two passes over the rows per product)
//   * the reduced system is solved by conjugate gradients preconditioned with the 6x6 (3x3) diagonal knot blocks of B, inverted per knot;
//   * every transposed product is a GATHER: one warp per knot sums the blocks of the rows whose window covers that knot, in the fixed order
//     of a row list sorted by first knot, and reduces across lanes with a fixed shuffle tree; per-landmark sums run over a row list sorted
//     by landmark.  No atomics anywhere: two runs give bit-identical results (tests/test_gn_device.py), and on several GPUs the only
//     exchange is the all-reduce of parameter-sized vectors between the calls below (kontiki_b200/gn.py);
//   * the CG scalars (alpha, beta, residual norm, iteration count, convergence flag) live in device memory: an iteration is four kernel
//     launches and no host round trip; the host looks at the flag every few iterations.
// Retraction (SE3: T exp(delta), uniform_se3_spline_trajectory.h:25-48; quaternions: ceres::EigenQuaternionParameterization; rho >= 0,
// static_rscamera_measurement.h:180) also happens on the device.
#pragma once

namespace {

constexpr int kGnMaxWin = 24;      // (group, window) pairs per spline

struct GnWinDev {            // one (group, window) of one spline, as the gather kernels see it
  const double* J;           // the group's packed rows (device order)
  const double* x;           // per-row vector that is transposed (u, or the residuals), nres per row
  const int* order;          // row indices sorted by the window's first knot
  const int* fk;             // first knot of order[j]
  const int* start;          // CSR over first knots: rows with first knot f are order[start[f] .. start[f+1])
  const int* partner_first;  // camera rows: first knot of the OTHER window of the same row on the same spline (row-indexed), else nullptr
  int partner_j_off;         // offset of that window's blocks in the row
  int role;                  // 0 plain, 1 reference window of a pair, 2 observation window of a pair (diagonal blocks: see k_gn_blocks)
  int j_off, width, nres, row_len;
};
struct GnWinList { GnWinDev w[kGnMaxWin]; int n; };

// d Plus / d delta at delta = 0.  SE3 knot (uniform_se3_spline_trajectory.h:25-48, Plus = T exp([upsilon; omega])): 7 x 6; SO3 knot
// (ceres::EigenQuaternionParameterization, Plus = q_delta q): 4 x 3.  Row-major.
__device__ __forceinline__ void gn_plus_se3(const double* k, double* P) {
  const double x = k[0], y = k[1], z = k[2], w = k[3];
#pragma unroll
  for (int i = 0; i < 42; ++i) P[i] = 0.0;
  P[0 * 6 + 3] = 0.5 * w;  P[0 * 6 + 4] = -0.5 * z; P[0 * 6 + 5] = 0.5 * y;
  P[1 * 6 + 3] = 0.5 * z;  P[1 * 6 + 4] = 0.5 * w;  P[1 * 6 + 5] = -0.5 * x;
  P[2 * 6 + 3] = -0.5 * y; P[2 * 6 + 4] = 0.5 * x;  P[2 * 6 + 5] = 0.5 * w;
  P[3 * 6 + 3] = -0.5 * x; P[3 * 6 + 4] = -0.5 * y; P[3 * 6 + 5] = -0.5 * z;
  const M3 R = quat_to_rot(x, y, z, w);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) P[(4 + i) * 6 + j] = R.a[3 * i + j];
}
__device__ __forceinline__ void gn_plus_so3(const double* q, double* P) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  P[0] = w;  P[1] = z;  P[2] = -y;
  P[3] = -z; P[4] = w;  P[5] = x;
  P[6] = y;  P[7] = -x; P[8] = w;
  P[9] = -x; P[10] = -y; P[11] = -z;
}
// kind 0: SE3 (width 7, local 6); 1: R3 (3, 3, identity); 2: SO3 (4, 3)
__global__ void k_gn_plus(const double* __restrict__ knots, int n, int kind, double* __restrict__ P) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  if (kind == 0) { double t[42]; gn_plus_se3(knots + (size_t)7 * k, t); for (int i = 0; i < 42; ++i) P[(size_t)42 * k + i] = t[i]; }
  else if (kind == 2) { double t[12]; gn_plus_so3(knots + (size_t)4 * k, t); for (int i = 0; i < 12; ++i) P[(size_t)12 * k + i] = t[i]; }
  else { for (int i = 0; i < 9; ++i) P[(size_t)9 * k + i] = (i % 4 == 0) ? 1.0 : 0.0; }
}

// va[k] = P_k v[k] (ambient from local), masked by `free` (0 for a locked spline)
__global__ void k_gn_to_ambient(const double* __restrict__ P, const double* __restrict__ v, int n, int width, int lw, double free_, double* __restrict__ va) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * width) return;
  const int k = i / width, c = i % width;
  double s = 0.0;
  for (int d = 0; d < lw; ++d) s += P[((size_t)k * width + c) * lw + d] * v[(size_t)k * lw + d];
  va[i] = free_ * s;
}

// u[row] = sum over windows / knots of J_block * va[knot]   (knot columns only: rho is eliminated), one thread per row
struct GnRowsArgs {
  RowWindows w; int n; const double* J; const int* idx[4]; const double* va[2]; double* u; const double* add; double add_scale;
};
__global__ void k_gn_rows_apply(const __grid_constant__ GnRowsArgs a) {      // __grid_constant__: the window tables are indexed in the constant bank, not copied to local memory
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const double* Jr = a.J + (size_t)i * a.w.row_len;
  double acc[3] = {0.0, 0.0, 0.0};
  for (int w = 0; w < a.w.nwin; ++w) {
    const int wd = a.w.width[w];
    const double* vv = a.va[a.w.col_off[w] ? 1 : 0] + (size_t)wd * a.idx[a.w.slot[w]][i];
    const double* Jw = Jr + a.w.j_off[w];
    for (int k = 0; k < 4; ++k)
      for (int r = 0; r < a.w.nres; ++r) {
        double s = 0.0;
        for (int c = 0; c < wd; ++c) s += Jw[(k * a.w.nres + r) * wd + c] * vv[k * wd + c];
        acc[r] += s;
      }
  }
  for (int r = 0; r < a.w.nres; ++r) a.u[(size_t)i * a.w.nres + r] = acc[r];
}

// Per-landmark sums over a group's rows (list sorted by landmark, CSR lm_start): out[l] (+)= sum_rows sum_r J_rho[row][r] * x[row][r]
// (x == nullptr: J_rho^2, the 1x1 block c_l).  One thread per landmark, fixed order.
__global__ void k_gn_lm_reduce(const double* __restrict__ J, int row_len, int rho_off, int nres, const double* __restrict__ x, const int* __restrict__ lm_order,
                               const int* __restrict__ lm_start, int n_lm, int accumulate, double* __restrict__ out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lm) return;
  double s = accumulate ? out[l] : 0.0;
  for (int j = lm_start[l]; j < lm_start[l + 1]; ++j) {
    const int row = lm_order[j];
    const double* jr = J + (size_t)row * row_len + rho_off;
    for (int r = 0; r < nres; ++r) s += jr[r] * (x ? x[(size_t)row * nres + r] : jr[r]);
  }
  out[l] = s;
}
// u[row] -= J_rho[row] * s[lm[row]]   (s = C^-1 t), or with `src` given: u[row] = src[row] - ...
__global__ void k_gn_rows_fix(const double* __restrict__ J, int row_len, int rho_off, int nres, const int* __restrict__ lm, const double* __restrict__ s, int n,
                              const double* __restrict__ src, double* __restrict__ u) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double sv = s[lm[i]];
  const double* jr = J + (size_t)i * row_len + rho_off;
  for (int r = 0; r < nres; ++r) u[(size_t)i * nres + r] = (src ? src[(size_t)i * nres + r] : u[(size_t)i * nres + r]) - jr[r] * sv;
}
// s[l] = free[l] ? t[l] / c[l] : 0
__global__ void k_gn_lm_scale(const double* __restrict__ t, const double* __restrict__ c, const unsigned char* __restrict__ locked, int n, double* __restrict__ s) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n) return;
  s[l] = (locked && locked[l]) || !(c[l] > 0.0) ? 0.0 : t[l] / c[l];
}

__device__ __forceinline__ double gn_warp_sum(double v) {      // fixed butterfly: every lane ends with the same, reproducible sum
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// y_loc[k] = P_k^T sum_{rows whose window covers knot k} J_block^T x_row : kGnWarpsPerKnot warps per knot, gather in the fixed order of the
// sorted row lists (lane l of warp w takes rows w*32 + l, + 32*kGnWarpsPerKnot, ...), fixed shuffle tree, fixed order over the warps.
constexpr int kGnWarpsPerKnot = 4;
template <int W>
__global__ void __launch_bounds__(32 * kGnWarpsPerKnot) k_gn_gather(const __grid_constant__ GnWinList L, int n_knots, int lw, const double* __restrict__ P, double* __restrict__ y) {
  __shared__ double sh[kGnWarpsPerKnot][8];
  const int k = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (k >= n_knots) return;
  double acc[W];
#pragma unroll
  for (int c = 0; c < W; ++c) acc[c] = 0.0;
  for (int wi = 0; wi < L.n; ++wi) {
    const GnWinDev& w = L.w[wi];
    const int f0 = max(k - 3, 0);
    for (int j = w.start[f0] + threadIdx.x; j < w.start[k + 1]; j += 32 * kGnWarpsPerKnot) {
      const int row = w.order[j], b = k - w.fk[j];
      const double* Jb = w.J + (size_t)row * w.row_len + w.j_off + b * w.nres * W;
      const double* xr = w.x + (size_t)row * w.nres;
      for (int r = 0; r < w.nres; ++r) {
        const double xv = xr[r];
#pragma unroll
        for (int c = 0; c < W; ++c) acc[c] += Jb[r * W + c] * xv;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < W; ++c) { acc[c] = gn_warp_sum(acc[c]); if (lane == 0) sh[wid][c] = acc[c]; }
  __syncthreads();
  if (threadIdx.x < lw) {
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < W; ++c) {
      double a = 0.0;
#pragma unroll
      for (int q = 0; q < kGnWarpsPerKnot; ++q) a += sh[q][c];
      s += P[((size_t)k * W + c) * lw + threadIdx.x] * a;
    }
    y[(size_t)k * lw + threadIdx.x] = s;
  }
}

// B_kk = P_k^T (sum_rows C_k^T C_k) P_k, C_k = the row's column block of knot k.  A camera row whose reference and observation windows both
// cover knot k has C_k = A + B (ONE parameter block, spline_base.h:391-394): the observation-window pass (role 2) uses A + B, the
// reference-window pass (role 1) skips such rows.  Output: n_knots x lw x lw (row-major), the exact diagonal block of J^T J in local coordinates.
template <int W>
__global__ void __launch_bounds__(32 * kGnWarpsPerKnot) k_gn_blocks(const __grid_constant__ GnWinList L, int n_knots, int lw, const double* __restrict__ P, double* __restrict__ Bd) {
  constexpr int NS = W * (W + 1) / 2;
  __shared__ double sh[kGnWarpsPerKnot][NS];
  __shared__ double Hs[NS];
  const int k = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (k >= n_knots) return;
  double H[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) H[i] = 0.0;
  for (int wi = 0; wi < L.n; ++wi) {
    const GnWinDev& w = L.w[wi];
    const int f0 = max(k - 3, 0);
    for (int j = w.start[f0] + threadIdx.x; j < w.start[k + 1]; j += 32 * kGnWarpsPerKnot) {
      const int row = w.order[j], b = k - w.fk[j];
      const double* Jb = w.J + (size_t)row * w.row_len + w.j_off + b * w.nres * W;
      int pb = -1;
      if (w.role != 0) { pb = k - w.partner_first[row]; if (pb < 0 || pb > 3) pb = -1; }
      if (w.role == 1 && pb >= 0) continue;
      const double* Jp = pb >= 0 ? w.J + (size_t)row * w.row_len + w.partner_j_off + pb * w.nres * W : nullptr;
      for (int r = 0; r < w.nres; ++r) {
        double cr[W];
#pragma unroll
        for (int c = 0; c < W; ++c) cr[c] = Jb[r * W + c] + (Jp ? Jp[r * W + c] : 0.0);
        int q = 0;
#pragma unroll
        for (int c = 0; c < W; ++c)
#pragma unroll
          for (int c2 = c; c2 < W; ++c2) H[q++] += cr[c] * cr[c2];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NS; ++i) { H[i] = gn_warp_sum(H[i]); if (lane == 0) sh[wid][i] = H[i]; }
  __syncthreads();
  if (threadIdx.x < NS) { double a = 0.0; for (int q = 0; q < kGnWarpsPerKnot; ++q) a += sh[q][threadIdx.x]; Hs[threadIdx.x] = a; }
  __syncthreads();
  // thread (d, e): (P^T H P)[d][e]
  if (threadIdx.x < lw * lw) {
    const int d = threadIdx.x / lw, e = threadIdx.x % lw;
    double s = 0.0;
    int q = 0;
    for (int c = 0; c < W; ++c)
      for (int c2 = c; c2 < W; ++c2) {
        const double h = Hs[q++];
        const double pcd = P[((size_t)k * W + c) * lw + d], pce = P[((size_t)k * W + c) * lw + e];
        const double p2d = P[((size_t)k * W + c2) * lw + d], p2e = P[((size_t)k * W + c2) * lw + e];
        s += (c == c2) ? h * pcd * pce : h * (pcd * p2e + p2d * pce);
      }
    Bd[(size_t)k * lw * lw + threadIdx.x] = s;
  }
}

// ---- conjugate gradients on the reduced (knot) system, scalars on the device ------------------------------------------------------
struct GnScal { double rz, pq, alpha, beta, rnorm2, bnorm2, tol2; int iter, done, max_iter, pad; double model_ur, model_uu, step2, cost; };

// M_k = (B_kk + D_k / radius)^-1 per knot (lw x lw SPD; D = clamp(diag B, 1e-6, 1e32): Ceres' LM scaling), Cholesky; also writes the
// damping diagonal D/radius.  A locked spline (free_ == 0) gets M = 0.
__global__ void k_gn_invert_blocks(const double* __restrict__ Bd, int n, int lw, double inv_radius, double free_, double* __restrict__ Minv, double* __restrict__ damp) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double A[36], Li[36];
  for (int i = 0; i < lw * lw; ++i) A[i] = Bd[(size_t)k * lw * lw + i];
  for (int d = 0; d < lw; ++d) {
    const double dd = fmin(fmax(A[d * lw + d], 1e-6), 1e32) * inv_radius;
    damp[(size_t)k * lw + d] = free_ * dd;
    A[d * lw + d] += dd;
  }
  // Cholesky A = L L^T, then A^-1 = L^-T L^-1
  bool ok = free_ != 0.0;
  for (int i = 0; i < lw && ok; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[i * lw + j];
      for (int m = 0; m < j; ++m) s -= A[i * lw + m] * A[j * lw + m];
      if (i == j) { if (!(s > 0.0)) { ok = false; break; } A[i * lw + i] = sqrt(s); }
      else A[i * lw + j] = s / A[j * lw + j];
    }
  if (!ok) { for (int i = 0; i < lw * lw; ++i) Minv[(size_t)k * lw * lw + i] = 0.0; return; }
  for (int c = 0; c < lw; ++c)            // Li = L^-1, column by column
    for (int i = 0; i < lw; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int m = 0; m < i; ++m) s -= A[i * lw + m] * Li[m * lw + c];
      Li[i * lw + c] = i < c ? 0.0 : s / A[i * lw + i];
    }
  for (int i = 0; i < lw; ++i)
    for (int j = 0; j < lw; ++j) {
      double s = 0.0;
      for (int m = max(i, j); m < lw; ++m) s += Li[m * lw + i] * Li[m * lw + j];
      Minv[(size_t)k * lw * lw + i * lw + j] = s;
    }
}

// deterministic block-wide sum (blockDim.x == 1024): fixed tree
__device__ double gn_block_sum(double v, double* sh) {
  v = gn_warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  double t = (threadIdx.x < 32) ? sh[threadIdx.x] : 0.0;
  if (wid == 0) t = gn_warp_sum(t);
  if (threadIdx.x == 0) sh[32] = t;
  __syncthreads();
  return sh[32];
}
struct GnVec { int n[2], lw[2]; const double* Minv[2]; const double* damp[2]; double* x[2]; double* r[2]; double* z[2]; double* p[2]; double* q[2]; const double* b[2]; };
__device__ __forceinline__ void gn_precond(const double* Minv, int lw, const double* r, double* z, int k) {
  for (int d = 0; d < lw; ++d) {
    double s = 0.0;
    for (int e = 0; e < lw; ++e) s += Minv[(size_t)k * lw * lw + d * lw + e] * r[(size_t)k * lw + e];
    z[(size_t)k * lw + d] = s;
  }
}
// The CG vector updates run on kGnCtas CTAs; every dot product is two-stage and reproducible: a CTA reduces its (fixed) slice with the fixed
// tree above into part[slot][cta], and the NEXT kernel of the chain lets every CTA add the kGnCtas partials in index order.
constexpr int kGnCtas = 64, kGnThreads = 256;
__device__ __forceinline__ double gn_cta_sum(double v, double* sh) {      // blockDim.x == kGnThreads (8 warps)
  v = gn_warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kGnThreads / 32; ++w) t += sh[w];
  return t;
}
__device__ __forceinline__ double gn_total(const double* part) {
  double t = 0.0;
  for (int i = 0; i < kGnCtas; ++i) t += part[i];
  return t;
}
// x = 0, r = b, z = M r, p = z; partials of r.z and |b|^2
__global__ void __launch_bounds__(kGnThreads) k_gn_pcg_init(const GnVec v, double* __restrict__ part) {
  __shared__ double sh[8];
  double rz = 0.0, bb = 0.0;
  for (int sp = 0; sp < 2; ++sp)
    for (int k = blockIdx.x * kGnThreads + threadIdx.x; k < v.n[sp]; k += kGnCtas * kGnThreads) {
      const int lw = v.lw[sp];
      for (int d = 0; d < lw; ++d) { const size_t i = (size_t)k * lw + d; v.x[sp][i] = 0.0; v.r[sp][i] = v.b[sp][i]; }
      gn_precond(v.Minv[sp], lw, v.r[sp], v.z[sp], k);
      for (int d = 0; d < lw; ++d) { const size_t i = (size_t)k * lw + d; v.p[sp][i] = v.z[sp][i]; rz += v.r[sp][i] * v.z[sp][i]; bb += v.b[sp][i] * v.b[sp][i]; }
    }
  rz = gn_cta_sum(rz, sh);
  bb = gn_cta_sum(bb, sh);
  if (threadIdx.x == 0) { part[blockIdx.x] = rz; part[kGnCtas + blockIdx.x] = bb; }
}
__global__ void k_gn_pcg_init_scal(const double* __restrict__ part, GnScal* s, double tol, int max_iter) {
  const double rz = gn_total(part), bb = gn_total(part + kGnCtas);
  s->rz = rz; s->bnorm2 = bb; s->rnorm2 = bb; s->tol2 = tol * tol; s->iter = 0; s->done = (bb == 0.0) ? 1 : 0; s->max_iter = max_iter;
}
// CG update, three kernels: (1) q += D p, partials of p.q; (2) alpha, x, r, z = M r, partials of r.z and r.r; (3) beta, p, scalars.
__global__ void __launch_bounds__(kGnThreads) k_gn_pcg_a(const GnVec v, const GnScal* __restrict__ s, double* __restrict__ part) {
  __shared__ double sh[8];
  if (s->done) return;
  double pq = 0.0;
  for (int sp = 0; sp < 2; ++sp)
    for (int i = blockIdx.x * kGnThreads + threadIdx.x; i < v.n[sp] * v.lw[sp]; i += kGnCtas * kGnThreads) { const double q = v.q[sp][i] + v.damp[sp][i] * v.p[sp][i]; v.q[sp][i] = q; pq += v.p[sp][i] * q; }
  pq = gn_cta_sum(pq, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = pq;
}
__global__ void __launch_bounds__(kGnThreads) k_gn_pcg_b(const GnVec v, const GnScal* __restrict__ s, const double* __restrict__ part, double* __restrict__ part2) {
  __shared__ double sh[8];
  if (s->done) return;
  const double alpha = s->rz / gn_total(part);
  double rz = 0.0, rr = 0.0;
  for (int sp = 0; sp < 2; ++sp)
    for (int k = blockIdx.x * kGnThreads + threadIdx.x; k < v.n[sp]; k += kGnCtas * kGnThreads) {
      const int lw = v.lw[sp];
      for (int d = 0; d < lw; ++d) { const size_t i = (size_t)k * lw + d; v.x[sp][i] += alpha * v.p[sp][i]; v.r[sp][i] -= alpha * v.q[sp][i]; }
      gn_precond(v.Minv[sp], lw, v.r[sp], v.z[sp], k);
      for (int d = 0; d < lw; ++d) { const size_t i = (size_t)k * lw + d; rz += v.r[sp][i] * v.z[sp][i]; rr += v.r[sp][i] * v.r[sp][i]; }
    }
  rz = gn_cta_sum(rz, sh);
  rr = gn_cta_sum(rr, sh);
  if (threadIdx.x == 0) { part2[blockIdx.x] = rz; part2[kGnCtas + blockIdx.x] = rr; }
}
__global__ void __launch_bounds__(kGnThreads) k_gn_pcg_c(const GnVec v, GnScal* s, const double* __restrict__ part, const double* __restrict__ part2, int* __restrict__ ticket) {
  if (s->done) return;
  const double rz = gn_total(part2), rr = gn_total(part2 + kGnCtas);
  const double beta = rz / s->rz;
  for (int sp = 0; sp < 2; ++sp)
    for (int i = blockIdx.x * kGnThreads + threadIdx.x; i < v.n[sp] * v.lw[sp]; i += kGnCtas * kGnThreads) v.p[sp][i] = v.z[sp][i] + beta * v.p[sp][i];
  // the LAST CTA to finish publishes the scalars (every CTA has read s->rz / s->done by then)
  __shared__ int last;
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); last = (atomicAdd(ticket, 1) == kGnCtas - 1); }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    const double pq = gn_total(part);
    s->pq = pq; s->alpha = s->rz / pq; s->beta = beta; s->rz = rz; s->rnorm2 = rr; s->iter += 1;
    if (rr <= s->tol2 * s->bnorm2 || s->iter >= s->max_iter) s->done = 1;
    *ticket = 0;
  }
}
// b = -y (right-hand side of the reduced system from the gathered gradient)
__global__ void k_gn_negate(const double* __restrict__ y, int n, double free_, double* __restrict__ b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) b[i] = -free_ * y[i];
}
// delta_rho[l] = -(g_rho[l] + t[l]) / c[l]  (t = E^T delta_k)
__global__ void k_gn_delta_rho(const double* __restrict__ g, const double* __restrict__ t, const double* __restrict__ c, const unsigned char* __restrict__ locked, int n,
                               double* __restrict__ d) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n) return;
  d[l] = ((locked && locked[l]) || !(c[l] > 0.0)) ? 0.0 : -(g[l] + t[l]) / c[l];
}
// rows: u += J_rho * delta_rho[lm]  (the full J delta), then partial sums of u.r and u.u for the model decrease  -(delta^T g + 1/2 delta^T J^T J delta)
__global__ void __launch_bounds__(256) k_gn_model_rows(const double* __restrict__ J, int row_len, int rho_off, int nres, const int* __restrict__ lm, const double* __restrict__ drho,
                                                       const double* __restrict__ r, double* __restrict__ u, int n, double* __restrict__ partial /* 2 per block */) {
  __shared__ double sh[2][8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double ur = 0.0, uu = 0.0;
  if (i < n) {
    const double dv = (rho_off >= 0) ? drho[lm[i]] : 0.0;
    for (int c = 0; c < nres; ++c) {
      double uv = u[(size_t)i * nres + c];
      if (rho_off >= 0) uv += J[(size_t)i * row_len + rho_off + c] * dv;
      ur += uv * r[(size_t)i * nres + c]; uu += uv * uv;
    }
  }
  ur = gn_warp_sum(ur); uu = gn_warp_sum(uu);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = ur; sh[1][threadIdx.x >> 5] = uu; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) { a += sh[0][w]; b += sh[1][w]; }
    partial[2 * blockIdx.x] = a; partial[2 * blockIdx.x + 1] = b;
  }
}
// cost partial sums: 1/2 sum rho(s) per row (estimator._cost: rows carry the corrected residual; huber == nullptr: no loss)
__global__ void __launch_bounds__(256) k_gn_cost_rows(const double* __restrict__ r, int nres, const double* __restrict__ huber, int n, double* __restrict__ partial) {
  __shared__ double sh[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double c = 0.0;
  if (i < n) {
    double s = 0.0;
    for (int k = 0; k < nres; ++k) s += r[(size_t)i * nres + k] * r[(size_t)i * nres + k];
    if (huber) { const double a2 = huber[i] * huber[i]; if (s > a2) s = 2.0 * s - a2; }
    c = 0.5 * s;
  }
  c = gn_warp_sum(c);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) { double a = 0.0; for (int w = 0; w < 8; ++w) a += sh[w]; partial[blockIdx.x] = a; }
}
// out[slot] (+)= sum of partial[0..n) (stride `stride`), one CTA, fixed order
__global__ void __launch_bounds__(1024) k_gn_sum_partials(const double* __restrict__ partial, int n, int stride, int accumulate, double* __restrict__ out) {
  __shared__ double sh[33];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[(size_t)i * stride];
  s = gn_block_sum(s, sh);
  if (threadIdx.x == 0) *out = (accumulate ? *out : 0.0) + s;
}
__global__ void __launch_bounds__(1024) k_gn_norm2(const double* __restrict__ a, int na, const double* __restrict__ b, int nb, const double* __restrict__ c, int nc, double* __restrict__ out) {
  __shared__ double sh[33];
  double s = 0.0;
  for (int i = threadIdx.x; i < na; i += blockDim.x) s += a[i] * a[i];
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s += b[i] * b[i];
  for (int i = threadIdx.x; i < nc; i += blockDim.x) s += c[i] * c[i];
  s = gn_block_sum(s, sh);
  if (threadIdx.x == 0) *out = s;
}

// ---- retraction on the device ------------------------------------------------------------------------------------------------------
// SE3 knot: T <- T exp([upsilon; omega]) with Sophus' exp (uniform_se3_spline_trajectory.h:25-36), quaternion re-normalised like SO3's product
__global__ void k_gn_retract_se3(const double* __restrict__ k_in, const double* __restrict__ delta, int n, double* __restrict__ k_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* k = k_in + (size_t)7 * i;
  const double* d = delta + (size_t)6 * i;
  const V3 ups = v3(d[0], d[1], d[2]), om = v3(d[3], d[4], d[5]);
  const double th2 = dot(om, om), th = sqrt(th2);
  double imag, real, cb, cc;
  if (th < 1e-10) { imag = 0.5 - th2 / 48.0; real = 1.0 - th2 / 8.0; cb = 0.5; cc = 1.0 / 6.0; }
  else { imag = sin(0.5 * th) / th; real = cos(0.5 * th); cb = (1.0 - cos(th)) / th2; cc = (th - sin(th)) / (th2 * th); }
  const V3 wu = cross(om, ups), wwu = cross(om, wu);
  const V3 t = ups + cb * wu + cc * wwu;                       // V upsilon
  const M3 R = quat_to_rot(k[0], k[1], k[2], k[3]);
  const V3 tt = R * t;
  const double bx = imag * om.x, by = imag * om.y, bz = imag * om.z, bw = real;
  const double ax = k[0], ay = k[1], az = k[2], aw = k[3];
  double qx = aw * bx + ax * bw + ay * bz - az * by, qy = aw * by + ay * bw + az * bx - ax * bz, qz = aw * bz + az * bw + ax * by - ay * bx,
         qw = aw * bw - ax * bx - ay * by - az * bz;
  const double inv = 1.0 / sqrt(qx * qx + qy * qy + qz * qz + qw * qw);
  double* o = k_out + (size_t)7 * i;
  o[0] = qx * inv; o[1] = qy * inv; o[2] = qz * inv; o[3] = qw * inv; o[4] = k[4] + tt.x; o[5] = k[5] + tt.y; o[6] = k[6] + tt.z;
}
// SO3 knot: q <- q_delta q, q_delta = (sin|d| d/|d|, cos|d|) (ceres::EigenQuaternionParameterization), re-normalised
__global__ void k_gn_retract_so3(const double* __restrict__ q_in, const double* __restrict__ delta, int n, double* __restrict__ q_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* b = q_in + (size_t)4 * i;
  const double* d = delta + (size_t)3 * i;
  const double nrm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  const double kf = nrm > 0.0 ? sin(nrm) / nrm : 1.0;
  const double ax = kf * d[0], ay = kf * d[1], az = kf * d[2], aw = cos(nrm);
  double qx = aw * b[0] + ax * b[3] + ay * b[2] - az * b[1], qy = aw * b[1] + ay * b[3] + az * b[0] - ax * b[2], qz = aw * b[2] + az * b[3] + ax * b[1] - ay * b[0],
         qw = aw * b[3] - ax * b[0] - ay * b[1] - az * b[2];
  const double inv = 1.0 / sqrt(qx * qx + qy * qy + qz * qz + qw * qw);
  double* o = q_out + (size_t)4 * i;
  o[0] = qx * inv; o[1] = qy * inv; o[2] = qz * inv; o[3] = qw * inv;
}
__global__ void k_gn_retract_add(const double* __restrict__ a, const double* __restrict__ d, int n, double lower, int use_lower, double* __restrict__ o) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = a[i] + d[i];
  o[i] = use_lower ? fmax(lower, v) : v;
}

}  // namespace
