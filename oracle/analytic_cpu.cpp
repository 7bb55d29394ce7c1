// oracle/analytic_cpu.cpp -- BENCH / TEST INFRASTRUCTURE ONLY (never loaded by kontiki_b200/).
//
// SURVEY.md section 8(d), "CPU reference timing": next to the faithful restatement of the reference (oracle_capi.cpp: one autodiff
// cost function per measurement, Dual<4> multipass exactly like ceres::DynamicAutoDiffCostFunction) the bench also reports an
// OPTIMISED CPU variant, so that the GPU / CPU ratio is not inflated by autodiff overhead alone.  This file is that variant: the
// product's own closed-form mathematics -- kontiki_b200/csrc/spline_math.cuh compiled for the host (the text the CUDA kernels run:
// knot-pair log hoisted into one prepass per evaluation point, analytic SE(3) Jacobians, packed rows) -- looped over the measurements
// with OpenMP on all host cores.  The landmark side is evaluated per measurement (no per-landmark hoist: that needs the host-side
// record table of the product).  It is a baseline that is REPORTED, not a fallback: nothing in the product can reach it.
#include <chrono>
#include <cstddef>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../kontiki_b200/csrc/spline_math.cuh"

using namespace kb;

extern "C" {

// One residual + Jacobian evaluation of (gyroscope, accelerometer, static-RS camera) rows on a UniformSE3SplineTrajectory.
// Arrays as in include/kontiki_b200.h (n_* may be 0); outputs r_* / J_* in the packed layouts (3 / 84 and 2 / 114 doubles per row).
// Returns the seconds spent inside the evaluation (knot packing + pair prepass + all rows); status[0] = rows that were out of range.
double kta_se3_evaluate(double t0, double dt, int n_knots, const double* knots7,
                        int n_gyro, const double* g_t, const double* g_y, const double* g_w,
                        int n_accel, const double* a_t, const double* a_y, const double* a_w,
                        int n_cam, const double* K, double readout, int rows, int model, const double* wc, double gamma,
                        const double* obs_uv, const double* obs_t0, const double* ref_uv, const double* ref_t0, const int* lm_idx,
                        const double* rho, const double* c_w, const double* huber_c, int nthreads,
                        double* g_r, double* g_J, double* a_r, double* a_J, double* c_r, double* c_J, int* status) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  SplineConst sp{t0, dt, n_knots, 0};
  ImuConst imu{0.0, 0.1, 1, {0.0, 0.0, 0.0}};
  CameraConst cam;
  std::memset(&cam, 0, sizeof(cam));
  if (n_cam > 0) {
    for (int i = 0; i < 9; ++i) cam.K[i] = K[i];
    const double* a = K; double k[9];                     // 3x3 inverse by cofactors, once (pinhole_camera.h:63-67 inverts K per call)
    k[0] = a[4] * a[8] - a[5] * a[7]; k[1] = a[2] * a[7] - a[1] * a[8]; k[2] = a[1] * a[5] - a[2] * a[4];
    k[3] = a[5] * a[6] - a[3] * a[8]; k[4] = a[0] * a[8] - a[2] * a[6]; k[5] = a[2] * a[3] - a[0] * a[5];
    k[6] = a[3] * a[7] - a[4] * a[6]; k[7] = a[1] * a[6] - a[0] * a[7]; k[8] = a[0] * a[4] - a[1] * a[3];
    const double det = a[0] * k[0] + a[1] * k[3] + a[2] * k[6];
    for (int i = 0; i < 9; ++i) cam.Kinv[i] = k[i] / det;
    const double q_id[4] = {0.0, 0.0, 0.0, 1.0}, p0[3] = {0.0, 0.0, 0.0};
    camera_set_pose(cam, q_id, p0);
    cam.time_offset = 0.0; cam.readout = readout; cam.row_delta = readout / (double)rows; cam.max_time_offset = 0.1; cam.time_offset_locked = 1;
    cam.rows = rows; cam.model = model; cam.wc[0] = wc ? wc[0] : 0.0; cam.wc[1] = wc ? wc[1] : 0.0; cam.gamma = gamma;
  }
  std::vector<double> knots8((size_t)n_knots * kKnotStride), pairs((size_t)n_knots * kPairStride, 0.0);
  long long bad = 0;
  const auto tic = std::chrono::steady_clock::now();
#pragma omp parallel
  {
#pragma omp for schedule(static)
    for (int i = 0; i < n_knots; ++i) {
      for (int c = 0; c < 7; ++c) knots8[(size_t)i * kKnotStride + c] = knots7[(size_t)i * 7 + c];
      knots8[(size_t)i * kKnotStride + 7] = 0.0;
    }
#pragma omp for schedule(static)
    for (int p = 1; p < n_knots; ++p)
      for (int dir = 0; dir <= 14; ++dir) pair_prepass_item(knots8.data(), p, dir, pairs.data());
#pragma omp for schedule(dynamic, 256) reduction(+ : bad) nowait
    for (int i = 0; i < n_gyro; ++i) {
      int i0;
      bad += imu_row(0, sp, imu, knots8.data(), pairs.data(), g_t[i], g_y + 3 * (size_t)i, g_w[i], g_r + 3 * (size_t)i, g_J + 84 * (size_t)i, &i0) != 0;
    }
#pragma omp for schedule(dynamic, 256) reduction(+ : bad) nowait
    for (int i = 0; i < n_accel; ++i) {
      int i0;
      bad += imu_row(1, sp, imu, knots8.data(), pairs.data(), a_t[i], a_y + 3 * (size_t)i, a_w[i], a_r + 3 * (size_t)i, a_J + 84 * (size_t)i, &i0) != 0;
    }
#pragma omp for schedule(dynamic, 256) reduction(+ : bad)
    for (int i = 0; i < n_cam; ++i) {
      double* row = c_J + (size_t)114 * i;
      Segment s0, s1;
      const int nseg = static_rs_segments(sp, cam, ref_t0[i], obs_t0[i], s0, s1);
      int ir; double ur;
      const int which = nseg == 0 ? -1 : locate_in_segments(nseg, s0, s1, static_rs_time(cam, ref_t0[i], ref_uv[2 * (size_t)i + 1]), sp.t0, sp.dt, ir, ur);
      if (which < 0) { ++bad; continue; }
      const Segment& sr = which == 0 ? s0 : s1;
      double rec[kRefStride];
      if (landmark_ref_row(sp, cam, knots8.data(), pairs.data(), ref_uv + 2 * (size_t)i, ref_t0[i], sr.start, sr.n, rho[lm_idx[i]], rec) != 0) { ++bad; continue; }
      ObsForward f;
      static_rs_row_locate(sp, cam, obs_uv + 2 * (size_t)i, obs_t0[i], ref_t0[i], f);
      static_rs_row_pose(knots8.data(), pairs.data(), f);
      ObsAdjoint adj;
      int i0r, i0o;
      if (static_rs_row_ref_half(cam, f, rec, obs_uv + 2 * (size_t)i, c_w[i], huber_c ? huber_c[i] : 0.0, c_r + 2 * (size_t)i, row, row + 112, &i0r, &i0o, adj) != 0) { ++bad; continue; }
      static_rs_row_obs_half(knots8.data(), pairs.data(), f, adj, row + 56);
    }
  }
  const auto toc = std::chrono::steady_clock::now();
  if (status) status[0] = (int)bad;
  return std::chrono::duration<double>(toc - tic).count();
}

}  // extern "C"
