// TEST HARNESS ONLY (built by tests/hostcheck.py into tests/_build/libhostcheck.so, never shipped, never loaded by
// kontiki_b200/).  Compiles the __host__ __device__ mathematics of kontiki_b200/csrc/spline_math.cuh for the host so
// that the CPU-only test suite can compare the exact text the CUDA kernels execute against the CPU oracle.
#include <cstddef>
#include <vector>

#include "../kontiki_b200/csrc/sensor_jac.cuh"
#include "../kontiki_b200/csrc/newton_math.cuh"

using namespace kb;

// camera model of the following hc_* camera calls (0 pinhole, 1 atan): set by hc_set_camera_model
static int g_model = 0; static double g_wc[2] = {0.0, 0.0}, g_gamma = 0.0;
static void finish_cam(CameraConst& cam, int rows) { cam.rows = rows; cam.model = g_model; cam.wc[0] = g_wc[0]; cam.wc[1] = g_wc[1]; cam.gamma = g_gamma; }

extern "C" {

void hc_set_camera_model(int model, double wc0, double wc1, double gamma) { g_model = model; g_wc[0] = wc0; g_wc[1] = wc1; g_gamma = gamma; }

// knots7: n x 7 (reference layout) -> padded records + pair records (what ktk_evaluate's pack + K0 do on the device)
void hc_prepass(const double* knots7, int n, double* knots8, double* pairs) {
  for (int i = 0; i < n; ++i) { for (int c = 0; c < 7; ++c) knots8[(size_t)i * kKnotStride + c] = knots7[(size_t)i * 7 + c]; knots8[(size_t)i * kKnotStride + 7] = 0.0; }
  for (int c = 0; c < kPairStride; ++c) pairs[c] = 0.0;
  for (int p = 1; p < n; ++p) for (int dir = 0; dir <= 14; ++dir) pair_prepass_item(knots8, p, dir, pairs);
}

void hc_imu(int which, double t0, double dt, int n_knots, int compat, double time_offset, double max_time_offset, int locked,
            const double* knots8, const double* pairs, int n, const double* t, const double* y, const double* w, double* r, double* J,
            int* i0, int* status) {
  SplineConst sp{t0, dt, n_knots, compat};
  ImuConst imu{time_offset, max_time_offset, locked, {0.0, 0.0, 0.0}};
  const int ny = which == 3 ? 4 : 3, nres = which == 3 ? 1 : 3, row = which == 3 ? 28 : 84;      // orientation rows: q (4), one residual, [4][1][7]
  for (int i = 0; i < n; ++i) {
    i0[i] = -1;
    status[i] = imu_row(which, sp, imu, knots8, pairs, t[i], y + ny * i, w[i], r + nres * i, J + (size_t)row * i, i0 + i);
  }
}

void hc_static_rs(double t0, double dt, int n_knots, const double* K, const double* Kinv, const double* q_ct, const double* p_ct,
                  double time_offset, double max_time_offset, int locked, double readout, int rows, const double* knots8,
                  const double* pairs, int n, const double* obs_uv, const double* obs_t0, const double* ref_uv, const double* ref_t0,
                  const int* lm_idx, const double* rho, const double* w, const double* huber_c, double* r, double* J, int* i0_ref,
                  int* i0_obs, int* status) {
  SplineConst sp{t0, dt, n_knots, 0};
  CameraConst cam;
  for (int i = 0; i < 9; ++i) { cam.K[i] = K[i]; cam.Kinv[i] = Kinv[i]; }
  camera_set_pose(cam, q_ct, p_ct);
  cam.time_offset = time_offset; cam.row_delta = readout / (double)rows; cam.readout = readout; cam.max_time_offset = max_time_offset;
  cam.time_offset_locked = locked;
  finish_cam(cam, rows);
  for (int i = 0; i < n; ++i) {
    i0_ref[i] = -1; i0_obs[i] = -1;
    double* row = J + (size_t)114 * i;
    double* rec = row + kRefInRow;                // the landmark record lands inside the row buffer, exactly as in the kernel
    // host part of the pipeline (ktk.cu upload_group): which segment of this residual holds the reference evaluation
    Segment s0, s1;
    const int nseg = static_rs_segments(sp, cam, ref_t0[i], obs_t0[i], s0, s1);
    int ir; double ur;
    const int which = nseg == 0 ? -1 : locate_in_segments(nseg, s0, s1, static_rs_time(cam, ref_t0[i], ref_uv[2 * i + 1]), sp.t0, sp.dt, ir, ur);
    if (which < 0) { status[i] = kStatusRange; continue; }
    const Segment& sr = which == 0 ? s0 : s1;
    // K_lm: landmark reference record
    status[i] = landmark_ref_row(sp, cam, knots8, pairs, ref_uv + 2 * i, ref_t0[i], sr.start, sr.n, rho[lm_idx[i]], rec);
    if (status[i] != 0) continue;
    // K_obs: the observation row
    ObsForward f;
    static_rs_row_locate(sp, cam, obs_uv + 2 * i, obs_t0[i], ref_t0[i], f);
    static_rs_row_pose(knots8, pairs, f);
    ObsAdjoint adj;
    status[i] = static_rs_row_ref_half(cam, f, rec, obs_uv + 2 * i, w[i], huber_c ? huber_c[i] : 0.0, r + 2 * i, row, row + 112, i0_ref + i, i0_obs + i, adj);
    if (status[i] == 0) static_rs_row_obs_half(knots8, pairs, f, adj, row + 56);
  }
}

static int g_newton_fast = 0;      // 2: every row in closed form, any number of evaluations (k_newton_rs_fast + k_newton_rs_rev); 0: every row in forward mode (k_newton_rs alone); 1: closed form for rows that stop after one or two evaluations (k_newton_rs_fast + k_newton_rs)
void hc_set_newton_fast(int on) { g_newton_fast = on; }
// NewtonRsCameraMeasurement rows: what k_landmark_ref + k_newton_rs do, one (row, direction) at a time.
// J: n x (58 + 14 W) packed [ref 4x(2x7) | obs W x(2x7) | rho 2]; kbase = first knot of the observation span.
void hc_newton_rs(double t0, double dt, int n_knots, const double* K, const double* Kinv, const double* q_ct, const double* p_ct,
                  double time_offset, double max_time_offset, int locked, double readout, int rows, const double* knots8,
                  const double* pairs, int n, const double* obs_uv, const double* obs_t0, const double* ref_uv, const double* ref_t0,
                  const int* lm_idx, const double* rho, const double* w, const double* huber_c, int W, double* r, double* J, int* i0_ref,
                  int* kbase_out, int* iterations, int* status) {
  SplineConst sp{t0, dt, n_knots, 0};
  CameraConst cam;
  for (int i = 0; i < 9; ++i) { cam.K[i] = K[i]; cam.Kinv[i] = Kinv[i]; }
  camera_set_pose(cam, q_ct, p_ct);
  cam.time_offset = time_offset; cam.row_delta = readout / (double)rows; cam.readout = readout; cam.max_time_offset = max_time_offset;
  cam.time_offset_locked = locked;
  finish_cam(cam, rows);
  const int row_len = 58 + 14 * W;
  for (int i = 0; i < n; ++i) {
    i0_ref[i] = -1; kbase_out[i] = -1; iterations[i] = 0;
    double rec[kRefStride];
    Segment s0, s1;
    const int nseg = static_rs_segments(sp, cam, ref_t0[i], obs_t0[i], s0, s1);
    int ir; double ur;
    const int which = nseg == 0 ? -1 : locate_in_segments(nseg, s0, s1, static_rs_time(cam, ref_t0[i], ref_uv[2 * i + 1]), sp.t0, sp.dt, ir, ur);
    if (which < 0) { status[i] = kStatusRange; continue; }
    const Segment& sr = which == 0 ? s0 : s1;
    status[i] = landmark_ref_row(sp, cam, knots8, pairs, ref_uv + 2 * i, ref_t0[i], sr.start, sr.n, rho[lm_idx[i]], rec);
    if (status[i] != 0) continue;
    const int kbase = newton_obs_window_base(sp, cam, obs_t0[i]);
    status[i] = (g_newton_fast == 2 ? newton_rs_row_reverse : g_newton_fast ? newton_rs_row_fast : newton_rs_row)(sp, cam, knots8, pairs, rec, obs_uv + 2 * i, obs_t0[i], ref_t0[i], kbase, W, w[i],
                                                                      huber_c ? huber_c[i] : 0.0, r + 2 * i, J + (size_t)row_len * i, iterations + i);
    i0_ref[i] = (int)rec[7]; kbase_out[i] = kbase;
  }
}
static int g_lifting_analytic = 0;      // 0: forward-mode directions (k_lifting_rs_fwd), 1: closed form (k_lifting_rs)
void hc_set_lifting_analytic(int on) { g_lifting_analytic = on; }
// LiftingRsCameraMeasurement rows: what k_landmark_ref + k_lifting_rs do.  J: n x (90 + 21 W) packed [ref 4x(3x7) | obs W x(3x7) | vt 3 | rho 3].
void hc_lifting_rs(double t0, double dt, int n_knots, const double* K, const double* Kinv, const double* q_ct, const double* p_ct,
                   double time_offset, double max_time_offset, int locked, double readout, int rows, const double* knots8,
                   const double* pairs, int n, const double* obs_uv, const double* obs_t0, const double* ref_uv, const double* ref_t0,
                   const int* lm_idx, const double* rho, const double* vt, const double* w, const double* huber_c, int W, double* r, double* J,
                   int* i0_ref, int* kbase_out, int* status) {
  SplineConst sp{t0, dt, n_knots, 0};
  CameraConst cam;
  for (int i = 0; i < 9; ++i) { cam.K[i] = K[i]; cam.Kinv[i] = Kinv[i]; }
  camera_set_pose(cam, q_ct, p_ct);
  cam.time_offset = time_offset; cam.row_delta = readout / (double)rows; cam.readout = readout; cam.max_time_offset = max_time_offset;
  cam.time_offset_locked = locked;
  finish_cam(cam, rows);
  const int row_len = 90 + 21 * W;
  for (int i = 0; i < n; ++i) {
    i0_ref[i] = -1; kbase_out[i] = -1;
    double rec[kRefStride];
    Segment s0, s1;
    const int nseg = static_rs_segments(sp, cam, ref_t0[i], obs_t0[i], s0, s1);
    int ir; double ur;
    const int which = nseg == 0 ? -1 : locate_in_segments(nseg, s0, s1, static_rs_time(cam, ref_t0[i], ref_uv[2 * i + 1]), sp.t0, sp.dt, ir, ur);
    if (which < 0) { status[i] = kStatusRange; continue; }
    const Segment& sr = which == 0 ? s0 : s1;
    status[i] = landmark_ref_row(sp, cam, knots8, pairs, ref_uv + 2 * i, ref_t0[i], sr.start, sr.n, rho[lm_idx[i]], rec);
    if (status[i] != 0) continue;
    const int kbase = newton_obs_window_base(sp, cam, obs_t0[i]);
    if (g_lifting_analytic)
      status[i] = lifting_rs_row_packed(sp, cam, knots8, pairs, rec, obs_uv + 2 * i, obs_t0[i], ref_t0[i], vt[i], kbase, W, w[i], huber_c ? huber_c[i] : 0.0,
                                        r + 3 * i, J + (size_t)row_len * i);
    else
      status[i] = lifting_rs_row(sp, cam, knots8, pairs, rec, obs_uv + 2 * i, obs_t0[i], ref_t0[i], vt[i], kbase, W, w[i], huber_c ? huber_c[i] : 0.0,
                                 r + 3 * i, J + (size_t)row_len * i);
    i0_ref[i] = (int)rec[7]; kbase_out[i] = kbase;
  }
}
// Sensor-block columns of NewtonRs (nres = 2, Js n x 16) / LiftingRs (nres = 3, Js n x 24) rows: what k_span_sensor does, one (row, column) at a
// time.  Layout [q_ct (nres x 4) | p_ct (nres x 3) | time offset (nres, zero: locked)], Huber-corrected like the knot columns.
void hc_span_sensor(int lifting, double t0, double dt, int n_knots, const double* K, const double* Kinv, const double* q_ct, const double* p_ct,
                    double time_offset, double max_time_offset, int locked, double readout, int rows, const double* knots8, const double* pairs,
                    int n, const double* obs_uv, const double* obs_t0, const double* ref_uv, const double* ref_t0, const int* lm_idx,
                    const double* rho, const double* vt, const double* w, const double* huber_c, int W, double* Js, int* status) {
  SplineConst sp{t0, dt, n_knots, 0};
  CameraConst cam;
  for (int i = 0; i < 9; ++i) { cam.K[i] = K[i]; cam.Kinv[i] = Kinv[i]; }
  camera_set_pose(cam, q_ct, p_ct);
  cam.time_offset = time_offset; cam.row_delta = readout / (double)rows; cam.readout = readout; cam.max_time_offset = max_time_offset;
  cam.time_offset_locked = locked;
  finish_cam(cam, rows);
  const int nres = lifting ? 3 : 2;
  for (int i = 0; i < n; ++i) {
    const int kbase = newton_obs_window_base(sp, cam, obs_t0[i]);
    status[i] = 0;
    for (int c = 0; c < 7 && status[i] == 0; ++c)
      status[i] = span_sensor_column(lifting != 0, sp, cam, knots8, pairs, ref_uv + 2 * i, ref_t0[i], rho[lm_idx[i]], obs_uv + 2 * i, obs_t0[i],
                                     lifting ? vt[i] : 0.0, kbase, W, w[i], huber_c ? huber_c[i] : 0.0, c, Js + (size_t)8 * nres * i);
  }
}
int hc_newton_window(double t0, double dt, double readout, double obs_t0) {
  SplineConst sp{t0, dt, 1 << 30, 0};
  CameraConst cam; cam.readout = readout; cam.time_offset_locked = 1; cam.max_time_offset = 0.0;
  return newton_obs_window_size(sp, cam, obs_t0);
}


// ---- split (R3 + SO3) trajectory ----------------------------------------------------------------------------------------
void hc_split_prepass(const double* vecs3, int n_r3, const double* quats, int n_so3, double* vecs4, double* pairs, int* status) {
  for (int i = 0; i < n_r3; ++i) { for (int c = 0; c < 3; ++c) vecs4[(size_t)i * kVecStride + c] = vecs3[(size_t)i * 3 + c]; vecs4[(size_t)i * kVecStride + 3] = 0.0; }
  for (int c = 0; c < kSo3PairStride; ++c) pairs[c] = 0.0;
  *status = 0;
  for (int p = 1; p < n_so3; ++p) for (int dir = 0; dir <= 8; ++dir) { const int st = so3_pair_prepass_item(quats, p, dir, pairs); if (st) *status = st; }
}

void hc_imu_split(int which, double t0_r3, double dt_r3, int n_r3, double t0_so3, double dt_so3, int n_so3, double time_offset,
                  double max_time_offset, int locked, const double* vecs4, const double* quats, const double* pairs, int n, const double* t,
                  const double* y, const double* w, double* r, double* J, int* i0_r3, int* i0_so3, int* status) {
  SplitConst sp{t0_r3, dt_r3, n_r3, t0_so3, dt_so3, n_so3};
  ImuConst imu{time_offset, max_time_offset, locked, {0.0, 0.0, 0.0}};
  const int row = which == 0 ? 48 : (which == 1 ? 84 : (which == 2 ? 36 : 16));
  const int ny = which == 3 ? 4 : 3, nres = which == 3 ? 1 : 3;
  for (int i = 0; i < n; ++i) {
    i0_r3[i] = -1; i0_so3[i] = -1;
    status[i] = imu_row_split(which, sp, imu, vecs4, quats, pairs, t[i], y + ny * i, w[i], r + nres * i, J + (size_t)row * i, i0_r3 + i, i0_so3 + i);
  }
}

void hc_static_rs_split(double t0_r3, double dt_r3, int n_r3, double t0_so3, double dt_so3, int n_so3, const double* K, const double* Kinv,
                        const double* q_ct, const double* p_ct, double time_offset, double max_time_offset, int locked, double readout, int rows,
                        const double* vecs4, const double* quats, const double* pairs, int n, const double* obs_uv, const double* obs_t0,
                        const double* ref_uv, const double* ref_t0, const int* lm_idx, const double* rho, const double* w, const double* huber_c,
                        double* r, double* J, int* idx /*[n][4]: ref R3, obs R3, ref SO3, obs SO3*/, int* status) {
  SplitConst sp{t0_r3, dt_r3, n_r3, t0_so3, dt_so3, n_so3};
  CameraConst cam;
  for (int i = 0; i < 9; ++i) { cam.K[i] = K[i]; cam.Kinv[i] = Kinv[i]; }
  camera_set_pose(cam, q_ct, p_ct);
  cam.time_offset = time_offset; cam.row_delta = readout / (double)rows; cam.readout = readout; cam.max_time_offset = max_time_offset;
  cam.time_offset_locked = locked;
  finish_cam(cam, rows);
  for (int i = 0; i < n; ++i) {
    for (int c = 0; c < 4; ++c) idx[4 * i + c] = -1;
    double* row = J + (size_t)114 * i;
    double* rec = row + kRefSplitInRow;
    const double tr = static_rs_time(cam, ref_t0[i], ref_uv[2 * i + 1]);
    Segment a0, a1, b0, b1; int ia, ib; double ua, ub;
    const int na = static_rs_segments_split(sp, cam, ref_t0[i], obs_t0[i], sp.t0_r3, sp.dt_r3, a0, a1);
    const int nb = static_rs_segments_split(sp, cam, ref_t0[i], obs_t0[i], sp.t0_so3, sp.dt_so3, b0, b1);
    const int wa = na == 0 ? -1 : locate_in_segments(na, a0, a1, tr, sp.t0_r3, sp.dt_r3, ia, ua);
    const int wb = nb == 0 ? -1 : locate_in_segments(nb, b0, b1, tr, sp.t0_so3, sp.dt_so3, ib, ub);
    if (wa < 0 || wb < 0) { status[i] = kStatusRange; continue; }
    const Segment& sa = wa == 0 ? a0 : a1; const Segment& sb = wb == 0 ? b0 : b1;
    status[i] = landmark_ref_row_split(sp, cam, vecs4, quats, pairs, ref_uv + 2 * i, ref_t0[i], sa.start, sa.n, sb.start, sb.n, rho[lm_idx[i]], rec);
    if (status[i] != 0) continue;
    ObsForwardSplit f;
    static_rs_row_forward_split(sp, cam, quats, pairs, obs_uv + 2 * i, obs_t0[i], ref_t0[i], f);
    status[i] = static_rs_row_finish_split(cam, vecs4, quats, pairs, f, rec, obs_uv + 2 * i, w[i], huber_c ? huber_c[i] : 0.0, r + 2 * i, row, row + 112, idx + 4 * i);
  }
}


// NewtonRs (lifting = 0) / LiftingRs (1) rows on a split trajectory: what k_landmark_ref_split + k_span_rs_split do, one (row, direction) at a time.
// J: n x span_split_row_len; idx [n][4] = ref R3 first knot, obs R3 span base, ref SO3 first knot, obs SO3 span base.
int hc_span_split_row_len(int lifting, int Wa, int Wb) { return span_split_row_len(lifting != 0, Wa, Wb); }
int hc_span_window(double t0, double dt, double readout, double obs_t0) { return span_window_size(t0, dt, readout, obs_t0); }
void hc_span_rs_split(int lifting, double t0_r3, double dt_r3, int n_r3, double t0_so3, double dt_so3, int n_so3, const double* K, const double* Kinv,
                      const double* q_ct, const double* p_ct, double time_offset, double max_time_offset, int locked, double readout, int rows,
                      const double* vecs4, const double* quats, const double* pairs, int n, const double* obs_uv, const double* obs_t0,
                      const double* ref_uv, const double* ref_t0, const int* lm_idx, const double* rho, const double* vt, const double* w,
                      const double* huber_c, int Wa, int Wb, double* r, double* J, int* idx, int* status) {
  SplitConst sp{t0_r3, dt_r3, n_r3, t0_so3, dt_so3, n_so3};
  CameraConst cam;
  for (int i = 0; i < 9; ++i) { cam.K[i] = K[i]; cam.Kinv[i] = Kinv[i]; }
  camera_set_pose(cam, q_ct, p_ct);
  cam.time_offset = time_offset; cam.row_delta = readout / (double)rows; cam.readout = readout; cam.max_time_offset = max_time_offset;
  cam.time_offset_locked = locked;
  finish_cam(cam, rows);
  const int nres = lifting ? 3 : 2, len = span_split_row_len(lifting != 0, Wa, Wb), ndir = span_split_ndir(lifting != 0, Wa, Wb);
  for (int i = 0; i < n; ++i) {
    for (int c = 0; c < 4; ++c) idx[4 * i + c] = -1;
    double rec[kRefSplitStride];
    const double tr = static_rs_time(cam, ref_t0[i], ref_uv[2 * i + 1]);
    Segment a0, a1, b0, b1; int ia, ib; double ua, ub;
    const int na = static_rs_segments_split(sp, cam, ref_t0[i], obs_t0[i], sp.t0_r3, sp.dt_r3, a0, a1);
    const int nb = static_rs_segments_split(sp, cam, ref_t0[i], obs_t0[i], sp.t0_so3, sp.dt_so3, b0, b1);
    const int wa = na == 0 ? -1 : locate_in_segments(na, a0, a1, tr, sp.t0_r3, sp.dt_r3, ia, ua);
    const int wb = nb == 0 ? -1 : locate_in_segments(nb, b0, b1, tr, sp.t0_so3, sp.dt_so3, ib, ub);
    if (wa < 0 || wb < 0) { status[i] = kStatusRange; continue; }
    const Segment& sa = wa == 0 ? a0 : a1; const Segment& sb = wb == 0 ? b0 : b1;
    status[i] = landmark_ref_row_split(sp, cam, vecs4, quats, pairs, ref_uv + 2 * i, ref_t0[i], sa.start, sa.n, sb.start, sb.n, rho[lm_idx[i]], rec);
    if (status[i] != 0) continue;
    const int ka = span_window_base(sp.t0_r3, sp.dt_r3, obs_t0[i]), kb = span_window_base(sp.t0_so3, sp.dt_so3, obs_t0[i]);
    for (int dir = 0; dir < ndir && status[i] == 0; ++dir)
      status[i] = span_split_column(lifting != 0, sp, cam, vecs4, quats, pairs, rec, obs_uv + 2 * i, obs_t0[i], ref_t0[i], lifting ? vt[i] : 0.0, ka, Wa, kb, Wb,
                                    w[i], huber_c ? huber_c[i] : 0.0, dir, r + nres * i, J + (size_t)len * i);
    idx[4 * i] = (int)rec[7]; idx[4 * i + 1] = ka; idx[4 * i + 2] = (int)rec[8]; idx[4 * i + 3] = kb;
  }
}
// ... and their sensor-block columns (relative pose of the camera): Js n x 16 / n x 24
void hc_span_sensor_split(int lifting, double t0_r3, double dt_r3, int n_r3, double t0_so3, double dt_so3, int n_so3, const double* K, const double* Kinv,
                          const double* q_ct, const double* p_ct, double time_offset, double max_time_offset, int locked, double readout, int rows,
                          const double* vecs4, const double* quats, const double* pairs, int n, const double* obs_uv, const double* obs_t0,
                          const double* ref_uv, const double* ref_t0, const int* lm_idx, const double* rho, const double* vt, const double* w,
                          const double* huber_c, int Wa, int Wb, double* Js, int* status) {
  SplitConst sp{t0_r3, dt_r3, n_r3, t0_so3, dt_so3, n_so3};
  CameraConst cam;
  for (int i = 0; i < 9; ++i) { cam.K[i] = K[i]; cam.Kinv[i] = Kinv[i]; }
  camera_set_pose(cam, q_ct, p_ct);
  cam.time_offset = time_offset; cam.row_delta = readout / (double)rows; cam.readout = readout; cam.max_time_offset = max_time_offset;
  cam.time_offset_locked = locked;
  finish_cam(cam, rows);
  const int nres = lifting ? 3 : 2;
  for (int i = 0; i < n; ++i) {
    const int ka = span_window_base(sp.t0_r3, sp.dt_r3, obs_t0[i]), kb = span_window_base(sp.t0_so3, sp.dt_so3, obs_t0[i]);
    status[i] = 0;
    for (int c = 0; c < 7 && status[i] == 0; ++c)
      status[i] = span_split_sensor_column(lifting != 0, sp, cam, vecs4, quats, pairs, ref_uv + 2 * i, ref_t0[i], rho[lm_idx[i]], obs_uv + 2 * i, obs_t0[i],
                                           lifting ? vt[i] : 0.0, ka, Wa, kb, Wb, w[i], huber_c ? huber_c[i] : 0.0, c, Js + (size_t)8 * nres * i);
  }
}
void hc_traj_eval_se3(double t0, double dt, int n_knots, int compat, const double* knots8, const double* pairs, int n, const double* t, double* out, int* status) {
  SplineConst sp{t0, dt, n_knots, compat};
  for (int i = 0; i < n; ++i) status[i] = traj_eval_se3(sp, knots8, pairs, t[i], out + 16 * i);
}
void hc_traj_eval_split(double t0_r3, double dt_r3, int n_r3, double t0_so3, double dt_so3, int n_so3, const double* vecs4, const double* quats,
                        const double* pairs, int n, const double* t, double* out, int* status) {
  SplitConst sp{t0_r3, dt_r3, n_r3, t0_so3, dt_so3, n_so3};
  for (int i = 0; i < n; ++i) status[i] = traj_eval_split(sp, vecs4, quats, pairs, t[i], out + 16 * i);
}


// ---- sensor-block Jacobians ---------------------------------------------------------------------------------------------
void hc_imu_time_offset_se3(int which, double t0, double dt, int n_knots, int compat, double time_offset, double max_time_offset, int locked,
                            const double* knots8, const double* pairs, int n, const double* t, const double* w, double* out, int* status) {
  SplineConst sp{t0, dt, n_knots, compat};
  ImuConst imu{time_offset, max_time_offset, locked, {0.0, 0.0, 0.0}};
  for (int i = 0; i < n; ++i) status[i] = imu_time_offset_jac_se3(which, sp, imu, knots8, pairs, t[i], w[i], out + 3 * i);
}
void hc_imu_time_offset_split(int which, double t0_r3, double dt_r3, int n_r3, double t0_so3, double dt_so3, int n_so3, double time_offset,
                              double max_time_offset, int locked, const double* vecs4, const double* quats, const double* pairs, int n, const double* t,
                              const double* w, double* out, int* status) {
  SplitConst sp{t0_r3, dt_r3, n_r3, t0_so3, dt_so3, n_so3};
  ImuConst imu{time_offset, max_time_offset, locked, {0.0, 0.0, 0.0}};
  for (int i = 0; i < n; ++i) status[i] = imu_time_offset_jac_split(which, sp, imu, vecs4, quats, pairs, t[i], w[i], out + 3 * i);
}
void hc_static_rs_sensor_se3(double t0, double dt, int n_knots, const double* K, const double* Kinv, const double* q_ct, const double* p_ct,
                             double time_offset, double max_time_offset, int locked, double readout, int rows, const double* knots8,
                             const double* pairs, int n, const double* obs_uv, const double* obs_t0, const double* ref_uv, const double* ref_t0,
                             const int* lm_idx, const double* rho, const double* w, const double* huber_c, double* out, int* status) {
  SplineConst sp{t0, dt, n_knots, 0};
  CameraConst cam;
  for (int i = 0; i < 9; ++i) { cam.K[i] = K[i]; cam.Kinv[i] = Kinv[i]; }
  camera_set_pose(cam, q_ct, p_ct);
  cam.time_offset = time_offset; cam.row_delta = readout / (double)rows; cam.readout = readout; cam.max_time_offset = max_time_offset;
  cam.time_offset_locked = locked;
  finish_cam(cam, rows);
  for (int i = 0; i < n; ++i)
    status[i] = static_rs_sensor_jac_se3(sp, cam, knots8, pairs, obs_uv + 2 * i, obs_t0[i], ref_uv + 2 * i, ref_t0[i], rho[lm_idx[i]], w[i],
                                         huber_c ? huber_c[i] : 0.0, out + 16 * i);
}

void hc_static_rs_sensor_split(double t0_r3, double dt_r3, int n_r3, double t0_so3, double dt_so3, int n_so3, const double* K, const double* Kinv, const double* q_ct,
                               const double* p_ct, double time_offset, double max_time_offset, int locked, double readout, int rows, const double* vecs4,
                               const double* quats, const double* pairs, int n, const double* obs_uv, const double* obs_t0, const double* ref_uv, const double* ref_t0,
                               const int* lm_idx, const double* rho, const double* w, const double* huber_c, double* out, int* status) {
  SplitConst sp{t0_r3, dt_r3, n_r3, t0_so3, dt_so3, n_so3};
  CameraConst cam;
  for (int i = 0; i < 9; ++i) { cam.K[i] = K[i]; cam.Kinv[i] = Kinv[i]; }
  camera_set_pose(cam, q_ct, p_ct);
  cam.time_offset = time_offset; cam.row_delta = readout / (double)rows; cam.readout = readout; cam.max_time_offset = max_time_offset;
  cam.time_offset_locked = locked;
  finish_cam(cam, rows);
  for (int i = 0; i < n; ++i)
    status[i] = static_rs_sensor_jac_split(sp, cam, vecs4, quats, pairs, obs_uv + 2 * i, obs_t0[i], ref_uv + 2 * i, ref_t0[i], rho[lm_idx[i]], w[i],
                                           huber_c ? huber_c[i] : 0.0, out + 16 * i);
}

void hc_se3_matrices(double t0, double dt, int n_knots, const double* knots8, const double* pairs, int n, const double* t, double* out, int* status) {
  SplineConst sp{t0, dt, n_knots, 0};
  for (int i = 0; i < n; ++i) status[i] = traj_eval_se3_matrices(sp, knots8, pairs, t[i], out + 48 * i);
}

}  // extern "C"
