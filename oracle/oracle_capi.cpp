// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle (loaded with ctypes by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the product).
//
// Each batch call mirrors what the reference does per measurement:
//   1. problem construction  = *Measurement::AddToEstimator (builds the segment metas and the
//      parameter-block list; measurements/gyroscope_measurement.h:75-105,
//      measurements/accelerometer_measurement.h:77-108, measurements/static_rscamera_measurement.h:130-198)
//   2. evaluation            = ceres::DynamicAutoDiffCostFunction<Residual>::Evaluate (double pass when no
//      Jacobian is requested, ceil(#active/4) Jet<double,4> passes otherwise)
// Only step 2 is timed (eval_seconds).  PARITY: values and Jacobians pinned by the reference's restated property tests and by the
// independent 60-digit transcription tests/mp_reference.py (see kontiki_ref.hpp header).
#include <chrono>
#include <cstring>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "kontiki_ref.hpp"

using namespace kto;

extern "C" {

typedef struct {
  int kind;                 // 0 SE3, 1 Split(R3+SO3), 2 R3, 3 SO3
  double dt_a, t0_a; int n_a; const double* knots_a;   // SE3 (7 doubles/knot: qx qy qz qw tx ty tz) or R3 (3)
  double dt_b, t0_b; int n_b; const double* knots_b;   // SO3 (4 doubles/knot: x y z w)
  int compat_zero_dB;       // 1 = reproduce the reference's Jet path for the SE3 accelerometer (dB == 0)
  int locked;               // trajectory locked -> knot blocks constant
} kto_traj;

typedef struct {
  double q_ct[4];           // x y z w
  double p_ct[3];
  double time_offset;
  double max_time_offset;
  int q_locked, p_locked, d_locked;
  int has_bias;             // ConstantBiasImu
  double abias[3], gbias[3];
  int abias_locked, gbias_locked;
} kto_sensor;

// model: 0 PinholeCamera, 1 AtanCamera (wc, gamma);  method: 0 StaticRsCameraMeasurement, 1 NewtonRsCameraMeasurement
typedef struct { double readout; int rows, cols; double K[9]; int model, method; double wc[2]; double gamma; } kto_camera;

enum { KTO_OK = 0, KTO_RANGE_ERROR = -1, KTO_RUNTIME_ERROR = -2, KTO_CAPACITY = -3 };

static thread_local std::string g_last_error;
const char* kto_last_error() { return g_last_error.c_str(); }
int kto_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"

namespace {

int knot_size(SplineKind k) { return k == kSE3 ? 7 : (k == kSO3 ? 4 : 3); }

struct TrajData {
  TrajKind kind; double dt_a, t0_a, dt_b, t0_b; int n_a, n_b; const double *ka, *kb; EvalOptions opt; bool locked;
  explicit TrajData(const kto_traj& t) : kind(TrajKind(t.kind)), dt_a(t.dt_a), t0_a(t.t0_a), dt_b(t.dt_b), t0_b(t.t0_b),
      n_a(t.n_a), n_b(t.n_b), ka(t.knots_a), kb(t.knots_b), locked(t.locked != 0) { opt.compat_zero_dB = t.compat_zero_dB != 0; }
  bool has_a() const { return kind != kTrajSO3; }
  bool has_b() const { return kind == kTrajSplit || kind == kTrajSO3; }
  SplineKind kind_a() const { return kind == kTrajSE3 ? kSE3 : kR3; }
  int size_a() const { return knot_size(kind_a()); }
  // trajectory_estimator.h / spline_base.h:47-55 / split_trajectory.h:60-66
  double MinTime() const {
    if (kind == kTrajSplit) return std::max(t0_a, t0_b);
    return has_a() ? t0_a : t0_b; }
  double MaxTime() const {
    auto mx = [](double t0, double dt, int n) { if (n < 4) throw std::range_error("Spline had too few control points"); return t0 + (size_t(n) - 3) * dt; };
    if (kind == kTrajSplit) return std::min(mx(t0_a, dt_a, n_a), mx(t0_b, dt_b, n_b));
    return has_a() ? mx(t0_a, dt_a, n_a) : mx(t0_b, dt_b, n_b); }
};

// One residual block as the reference's AddToEstimator would have registered it.
struct Block {
  TrajMeta meta;                       // segment metas (residual->trajectory_meta)
  std::vector<int> ids_a, ids_b;       // knot indices, in parameter-block order
  std::vector<const double*> params;   // parameter block pointers, reference order
  std::vector<int> sizes;
  std::vector<char> constant;
};

// TrajectoryEstimator::AddTrajectoryForTimes (trajectory_estimator.h:75-81) + {Spline,Split}Entity::AddToProblem
void add_trajectory(const TrajData& td, const std::vector<std::pair<double, double>>& times, Block& b) {
  check_time_spans(times, td.MinTime(), td.MaxTime());
  b.meta.kind = td.kind; b.meta.opt = td.opt;
  if (td.has_a()) {
    spline_add_to_problem(td.dt_a, td.t0_a, times, b.meta.a, b.ids_a);
    for (int id : b.ids_a) {
      if (id < 0 || id >= td.n_a) throw std::range_error("knot index out of range");   // vector::at in dynamic_pstore
      b.params.push_back(td.ka + size_t(id) * td.size_a()); b.sizes.push_back(td.size_a()); b.constant.push_back(td.locked);
    }
  }
  if (td.has_b()) {
    spline_add_to_problem(td.dt_b, td.t0_b, times, b.meta.b, b.ids_b);
    for (int id : b.ids_b) {
      if (id < 0 || id >= td.n_b) throw std::range_error("knot index out of range");
      b.params.push_back(td.kb + size_t(id) * 4); b.sizes.push_back(4); b.constant.push_back(td.locked);
    }
  }
}
// SensorEntity::AddToProblem (sensors/sensors.h:135-165) (+ constant_bias_imu.h:100-119: abias then gbias)
void add_sensor(const kto_sensor& s, Block& b) {
  b.params.push_back(s.q_ct); b.sizes.push_back(4); b.constant.push_back(s.q_locked != 0);
  b.params.push_back(s.p_ct); b.sizes.push_back(3); b.constant.push_back(s.p_locked != 0);
  b.params.push_back(&s.time_offset); b.sizes.push_back(1); b.constant.push_back(s.d_locked != 0);
  if (s.has_bias) {
    b.params.push_back(s.abias); b.sizes.push_back(3); b.constant.push_back(s.abias_locked != 0);
    b.params.push_back(s.gbias); b.sizes.push_back(3); b.constant.push_back(s.gbias_locked != 0);
  }
}

struct ImuFunctor {   // measurements/gyroscope_measurement.h:54-72 / accelerometer_measurement.h:56-74
  const Block* blk; int which; double t, weight; const double* y; bool has_bias;
  template <class T> bool operator()(T const* const* params, T* residual) const {
    size_t offset = 0;
    const TrajectoryView<T> trajectory(blk->meta, &params[offset]);
    offset += blk->meta.NumParameters();
    SensorView<T> imu; imu.params = &params[offset]; imu.has_bias = has_bias;
    if (which == 0) gyro_error<T>(weight, y, t, imu, trajectory, residual);
    else if (which == 2) position_error<T>(weight, y, t, trajectory, residual);      // PositionMeasurement: the sensor blocks are inert
    else if (which == 3) orientation_error<T>(y, t, trajectory, residual);           // OrientationMeasurement: y = q (x,y,z,w), 1 residual
    else accel_error<T>(weight, y, t, imu, trajectory, residual);
    return true;
  }
};

struct StaticRsFunctor {   // measurements/static_rscamera_measurement.h:108-127 / newton_rscamera_measurement.h:183-199
  const Block* blk; CameraMeta cm; double weight; const double *ref_uv, *obs_uv; double ref_t0, obs_t0; int method;
  template <class T> bool operator()(T const* const* params, T* residual) const {
    size_t offset = 0;
    const TrajectoryView<T> trajectory(blk->meta, &params[offset]);
    offset += blk->meta.NumParameters();
    SensorView<T> camera; camera.params = &params[offset]; camera.has_bias = false;
    offset += 3;
    T inverse_depth = params[offset][0];
    T y_hat[2];
    if (method == 1) reproject_newton<T>(cm, ref_uv, ref_t0, obs_uv, obs_t0, inverse_depth, trajectory, trajectory, camera, y_hat);
    else reproject_static<T>(cm, ref_uv, ref_t0, obs_uv, obs_t0, inverse_depth, trajectory, trajectory, camera, y_hat);
    residual[0] = T(weight) * (T(obs_uv[0]) - y_hat[0]);   // :89-94
    residual[1] = T(weight) * (T(obs_uv[1]) - y_hat[1]);
    return true;
  }
};

struct LiftingRsFunctor {   // measurements/lifting_rscamera_measurement.h:105-149: blocks [trajectory | camera (3) | vt | rho], 3 residuals
  const Block* blk; CameraMeta cm; double weight; const double *ref_uv, *obs_uv; double ref_t0, obs_t0, vt_orig;
  template <class T> bool operator()(T const* const* params, T* residual) const {
    size_t offset = 0;
    const TrajectoryView<T> trajectory(blk->meta, &params[offset]);
    offset += blk->meta.NumParameters();
    SensorView<T> camera; camera.params = &params[offset]; camera.has_bias = false;
    offset += 3;
    T vt = params[offset][0];
    offset += 1;
    T inverse_depth = params[offset][0];
    T y_hat[2];
    reproject_lifting<T>(cm, ref_uv, ref_t0, obs_t0, vt, inverse_depth, trajectory, camera, y_hat);
    residual[0] = T(weight) * (T(obs_uv[0]) - y_hat[0]);                       // :105-110 projection error (pixels)
    residual[1] = T(weight) * (T(obs_uv[1]) - y_hat[1]);
    residual[2] = T(weight) * (T(double(cm.rows)) * (vt - T(vt_orig)));        // :112-113 timing error (pixels)
    return true;
  }
};

// SplineView::Evaluate segment choice + CalculateIndexAndInterpolationAmount -> GLOBAL index of the first
// active knot (spline_base.h:188-202, 148-152); returns -1 if no segment holds t.
int locate_knot(const SplineMeta& meta, const std::vector<int>& ids, double t) {
  size_t offset = 0;
  for (auto& seg : meta.segments) {
    if (t >= seg.MinTime() && t < seg.MaxTime()) {
      double s = (t - seg.t0) / seg.dt; int i0 = int(std::floor(s));
      if (i0 < 0 || size_t(i0) > seg.n - 4) return -1;
      return ids[offset + i0];
    }
    offset += seg.n;
  }
  return -1;
}

template <class F> int guarded(F&& f) {
  try { f(); return KTO_OK; }
  catch (const std::range_error& e) { g_last_error = e.what(); return KTO_RANGE_ERROR; }
  catch (const std::exception& e) { g_last_error = e.what(); return KTO_RUNTIME_ERROR; }
}

// Evaluate one block; jac_mode 0 none, 1 non-constant blocks (what Ceres asks for), 2 all blocks.
template <class Functor>
void evaluate_block(const Functor& f, const Block& b, int num_res, int jac_mode, double* r, std::vector<std::vector<double>>& jac) {
  if (jac_mode == 0) { autodiff_evaluate(f, b.sizes, b.params.data(), num_res, r, nullptr); return; }
  const size_t nb = b.sizes.size();
  jac.resize(nb);
  std::vector<double*> jp(nb, nullptr);
  for (size_t i = 0; i < nb; ++i) if (jac_mode == 2 || !b.constant[i]) { jac[i].assign(size_t(num_res) * b.sizes[i], 0.0); jp[i] = jac[i].data(); } else jac[i].clear();
  autodiff_evaluate(f, b.sizes, b.params.data(), num_res, r, jp.data());
}

void copy_block(const std::vector<double>& src, double* dst, size_t count) {
  if (src.empty()) std::memset(dst, 0, count * sizeof(double)); else std::memcpy(dst, src.data(), count * sizeof(double));
}

}  // namespace

extern "C" {

// Whole-spline evaluation with T=double, as Python's traj.position(t) etc. do.
// out arrays may be NULL.  quat is (x,y,z,w).  status[i] per time.
int kto_traj_evaluate(const kto_traj* tr, int n, const double* t, int flags, double* pos, double* vel, double* acc,
                      double* quat, double* angvel, int* status) {
  TrajData td(*tr); int worst = KTO_OK;
  for (int i = 0; i < n; ++i) {
    int st = guarded([&] {
      Block b; b.meta.kind = td.kind; b.meta.opt = td.opt;
      if (td.has_a()) { SplineSegmentMeta sm; sm.t0 = td.t0_a; sm.dt = td.dt_a; sm.n = td.n_a; b.meta.a.segments.push_back(sm); for (int k = 0; k < td.n_a; ++k) b.params.push_back(td.ka + size_t(k) * td.size_a()); }
      if (td.has_b()) { SplineSegmentMeta sm; sm.t0 = td.t0_b; sm.dt = td.dt_b; sm.n = td.n_b; b.meta.b.segments.push_back(sm); for (int k = 0; k < td.n_b; ++k) b.params.push_back(td.kb + size_t(k) * 4); }
      TrajectoryView<double> view(b.meta, b.params.data());
      auto res = view.Evaluate(t[i], flags);
      if (pos) { pos[3 * i] = res->position.x; pos[3 * i + 1] = res->position.y; pos[3 * i + 2] = res->position.z; }
      if (vel) { vel[3 * i] = res->velocity.x; vel[3 * i + 1] = res->velocity.y; vel[3 * i + 2] = res->velocity.z; }
      if (acc) { acc[3 * i] = res->acceleration.x; acc[3 * i + 1] = res->acceleration.y; acc[3 * i + 2] = res->acceleration.z; }
      if (quat) { quat[4 * i] = res->orientation.x; quat[4 * i + 1] = res->orientation.y; quat[4 * i + 2] = res->orientation.z; quat[4 * i + 3] = res->orientation.w; }
      if (angvel) { angvel[3 * i] = res->angular_velocity.x; angvel[3 * i + 1] = res->angular_velocity.y; angvel[3 * i + 2] = res->angular_velocity.z; }
    });
    if (status) status[i] = st;
    if (st != KTO_OK) worst = st;
  }
  return worst;
}

// UniformSE3SplineTrajectory.evaluate(t) -> (P, P', P'') 4x4 row-major (py_uniform_se3_spline_trajectory.cc:53-60, flags 0xff)
int kto_se3_evaluate_matrices(const kto_traj* tr, int n, const double* t, double* P, double* Pp, double* Pb, int* status) {
  TrajData td(*tr); int worst = KTO_OK;
  for (int i = 0; i < n; ++i) {
    int st = guarded([&] {
      std::vector<const double*> params; for (int k = 0; k < td.n_a; ++k) params.push_back(td.ka + size_t(k) * 7);
      SegmentView<double> sv; sv.meta.t0 = td.t0_a; sv.meta.dt = td.dt_a; sv.meta.n = td.n_a; sv.params = params.data(); sv.kind = kSE3; sv.opt = td.opt;
      SE3<double> Pse3; Mat4<double> m1 = mat4_zero<double>(), m2 = mat4_zero<double>();
      sv.EvaluateSplineSE3(t[i], 0xff, Pse3, m1, m2);
      Mat4<double> m0 = se3_matrix(Pse3);
      for (int a = 0; a < 4; ++a) for (int c = 0; c < 4; ++c) { P[16 * i + 4 * a + c] = m0.m[a][c]; Pp[16 * i + 4 * a + c] = m1.m[a][c]; Pb[16 * i + 4 * a + c] = m2.m[a][c]; }
    });
    if (status) status[i] = st;
    if (st != KTO_OK) worst = st;
  }
  return worst;
}

// Gyroscope (which=0) / accelerometer (which=1) / PositionMeasurement (which=2, position_measurement.h) / OrientationMeasurement
// (which=3, orientation_measurement.h: y has 4 doubles per row (x,y,z,w), ONE residual per row: r[n], blocks 1 x size) residual blocks.
//   r[3n]; cap_a / cap_b = capacity (knots per measurement) of ids_a/Ja and ids_b/Jb;
//   ids_a[n*cap_a] (-1 padded), Ja[n*cap_a*3*size_a] row-major 3 x size blocks; same for b (SO3, size 4);
//   Js[n*42]: q_ct 3x4 | p_ct 3x3 | d 3x1 | abias 3x3 | gbias 3x3 ; i0_a/i0_b: global index of the first ACTIVE knot.
//   jac_mode: 0 none, 1 non-constant blocks only (Ceres behaviour; used for timing), 2 every block.
int kto_imu_residuals(const kto_traj* tr, const kto_sensor* imu, int which, int n, const double* t, const double* y,
                      const double* weight, int jac_mode, int nthreads, double* r, int cap_a, int* ids_a, double* Ja,
                      int cap_b, int* ids_b, double* Jb, double* Js, int* i0_a, int* i0_b, int* status, double* eval_seconds) {
  TrajData td(*tr);
  std::vector<Block> blocks(n);
  std::vector<int> st(n, KTO_OK);
  std::string first_err;
  // --- problem construction (untimed)
  for (int i = 0; i < n; ++i) {
    st[i] = guarded([&] {
      double tmin, tmax;   // gyroscope_measurement.h:82-91
      if (imu->d_locked) { tmin = t[i]; tmax = t[i]; } else { tmin = t[i] - imu->max_time_offset; tmax = t[i] + imu->max_time_offset; }
      add_trajectory(td, {{tmin, tmax}}, blocks[i]);
      add_sensor(*imu, blocks[i]);
      if (int(blocks[i].ids_a.size()) > cap_a && ids_a) throw std::length_error("cap_a too small");
      if (int(blocks[i].ids_b.size()) > cap_b && ids_b) throw std::length_error("cap_b too small");
    });
    if (st[i] != KTO_OK && first_err.empty()) first_err = g_last_error;
  }
  const int sa = td.has_a() ? td.size_a() : 0;
  const int nres = which == 3 ? 1 : 3, ny = which == 3 ? 4 : 3;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  auto tic = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; ++i) {
    if (st[i] != KTO_OK) continue;
    const Block& b = blocks[i];
    ImuFunctor f{&b, which, t[i], weight ? weight[i] : 1.0, y + ny * i, imu->has_bias != 0};
    std::vector<std::vector<double>> jac;
    st[i] = guarded([&] { evaluate_block(f, b, nres, jac_mode, r + nres * i, jac); });
    if (st[i] != KTO_OK) continue;
    const double te = t[i] + imu->time_offset;
    if (i0_a) i0_a[i] = td.has_a() ? locate_knot(b.meta.a, b.ids_a, te) : -1;
    if (i0_b) i0_b[i] = td.has_b() ? locate_knot(b.meta.b, b.ids_b, te) : -1;
    if (ids_a) for (int k = 0; k < cap_a; ++k) ids_a[size_t(i) * cap_a + k] = k < int(b.ids_a.size()) ? b.ids_a[k] : -1;
    if (ids_b) for (int k = 0; k < cap_b; ++k) ids_b[size_t(i) * cap_b + k] = k < int(b.ids_b.size()) ? b.ids_b[k] : -1;
    if (jac_mode == 0) continue;
    size_t pb = 0;
    if (Ja) { std::memset(Ja + size_t(i) * cap_a * nres * sa, 0, sizeof(double) * cap_a * nres * sa);
      for (size_t k = 0; k < b.ids_a.size(); ++k) copy_block(jac[pb + k], Ja + (size_t(i) * cap_a + k) * nres * sa, nres * sa); }
    pb += b.ids_a.size();
    if (Jb) { std::memset(Jb + size_t(i) * cap_b * nres * 4, 0, sizeof(double) * cap_b * nres * 4);
      for (size_t k = 0; k < b.ids_b.size(); ++k) copy_block(jac[pb + k], Jb + (size_t(i) * cap_b + k) * nres * 4, nres * 4); }
    pb += b.ids_b.size();
    if (Js && nres == 3) {
      double* d = Js + size_t(i) * 42; std::memset(d, 0, 42 * sizeof(double));
      copy_block(jac[pb], d, 12); copy_block(jac[pb + 1], d + 12, 9); copy_block(jac[pb + 2], d + 21, 3);
      if (imu->has_bias) { copy_block(jac[pb + 3], d + 24, 9); copy_block(jac[pb + 4], d + 33, 9); }
    }
  }
  auto toc = std::chrono::steady_clock::now();
  if (eval_seconds) *eval_seconds = std::chrono::duration<double>(toc - tic).count();
  int worst = KTO_OK;
  for (int i = 0; i < n; ++i) { if (status) status[i] = st[i]; if (st[i] != KTO_OK) worst = st[i]; }
  if (worst != KTO_OK && !first_err.empty()) g_last_error = first_err;
  return worst;
}

// StaticRsCameraMeasurement residual blocks (SE3 or Split trajectory).
//   per measurement: obs_uv[2], obs_t0, ref_uv[2], ref_t0, lm_idx -> rho[lm_idx], weight
//   r[2n]; ids_a/Ja (2 x size_a blocks), ids_b/Jb (2x4), Js[n*16] (q_ct 2x4 | p_ct 2x3 | d 2x1 | pad), Jrho[2n]
//   i0_ref_a, i0_obs_a, i0_ref_b, i0_obs_b: global first-active-knot indices of the two evaluations.
int kto_static_rs_residuals(const kto_traj* tr, const kto_sensor* cam, const kto_camera* cmeta, int n, const double* obs_uv,
                            const double* obs_t0, const double* ref_uv, const double* ref_t0, const int* lm_idx, const double* rho,
                            const char* lm_locked, const double* weight, int jac_mode, int nthreads, double* r, int cap_a, int* ids_a,
                            double* Ja, int cap_b, int* ids_b, double* Jb, double* Js, double* Jrho, int* i0_ref_a, int* i0_obs_a,
                            int* i0_ref_b, int* i0_obs_b, int* status, double* eval_seconds) {
  TrajData td(*tr);
  CameraMeta cm; cm.readout = cmeta->readout; cm.rows = cmeta->rows; cm.cols = cmeta->cols;
  for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) cm.K[a][c] = cmeta->K[3 * a + c];
  cm.model = cmeta->model; cm.wc[0] = cmeta->wc[0]; cm.wc[1] = cmeta->wc[1]; cm.gamma = cmeta->gamma;
  const int method = cmeta->method;
  std::vector<Block> blocks(n);
  std::vector<int> st(n, KTO_OK);
  std::string first_err;
  for (int i = 0; i < n; ++i) {
    st[i] = guarded([&] {
      // static_rscamera_measurement.h:137-166
      double t1, t2;
      if (ref_t0[i] <= obs_t0[i]) { t1 = ref_t0[i]; t2 = obs_t0[i]; } else { t1 = obs_t0[i]; t2 = ref_t0[i]; }
      if (!cam->d_locked) { t1 -= cam->max_time_offset; t2 += cam->max_time_offset; }
      const double margin = 1e-3;
      add_trajectory(td, {{t1 - margin, t1 + cm.readout + margin}, {t2 - margin, t2 + cm.readout + margin}}, blocks[i]);
      add_sensor(*cam, blocks[i]);
      blocks[i].params.push_back(rho + lm_idx[i]); blocks[i].sizes.push_back(1);
      blocks[i].constant.push_back(lm_locked ? lm_locked[lm_idx[i]] != 0 : 0);
      if (int(blocks[i].ids_a.size()) > cap_a && ids_a) throw std::length_error("cap_a too small");
      if (int(blocks[i].ids_b.size()) > cap_b && ids_b) throw std::length_error("cap_b too small");
    });
    if (st[i] != KTO_OK && first_err.empty()) first_err = g_last_error;
  }
  const int sa = td.has_a() ? td.size_a() : 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  auto tic = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; ++i) {
    if (st[i] != KTO_OK) continue;
    const Block& b = blocks[i];
    StaticRsFunctor f{&b, cm, weight ? weight[i] : 1.0, ref_uv + 2 * i, obs_uv + 2 * i, ref_t0[i], obs_t0[i], method};
    std::vector<std::vector<double>> jac;
    st[i] = guarded([&] { evaluate_block(f, b, 2, jac_mode, r + 2 * i, jac); });
    if (st[i] != KTO_OK) continue;
    const double row_delta = cm.readout / double(cm.rows);
    const double t_ref = ref_t0[i] + cam->time_offset + ref_uv[2 * i + 1] * row_delta;
    // NewtonRs evaluates at several row times inside [t0_obs, t0_obs + readout]: report the window of the lower bound
    const double t_obs = obs_t0[i] + cam->time_offset + (method == 1 ? 0.0 : obs_uv[2 * i + 1] * row_delta);
    if (i0_ref_a) i0_ref_a[i] = td.has_a() ? locate_knot(b.meta.a, b.ids_a, t_ref) : -1;
    if (i0_obs_a) i0_obs_a[i] = td.has_a() ? locate_knot(b.meta.a, b.ids_a, t_obs) : -1;
    if (i0_ref_b) i0_ref_b[i] = td.has_b() ? locate_knot(b.meta.b, b.ids_b, t_ref) : -1;
    if (i0_obs_b) i0_obs_b[i] = td.has_b() ? locate_knot(b.meta.b, b.ids_b, t_obs) : -1;
    if (ids_a) for (int k = 0; k < cap_a; ++k) ids_a[size_t(i) * cap_a + k] = k < int(b.ids_a.size()) ? b.ids_a[k] : -1;
    if (ids_b) for (int k = 0; k < cap_b; ++k) ids_b[size_t(i) * cap_b + k] = k < int(b.ids_b.size()) ? b.ids_b[k] : -1;
    if (jac_mode == 0) continue;
    size_t pb = 0;
    if (Ja) { std::memset(Ja + size_t(i) * cap_a * 2 * sa, 0, sizeof(double) * cap_a * 2 * sa);
      for (size_t k = 0; k < b.ids_a.size(); ++k) copy_block(jac[pb + k], Ja + (size_t(i) * cap_a + k) * 2 * sa, 2 * sa); }
    pb += b.ids_a.size();
    if (Jb) { std::memset(Jb + size_t(i) * cap_b * 8, 0, sizeof(double) * cap_b * 8);
      for (size_t k = 0; k < b.ids_b.size(); ++k) copy_block(jac[pb + k], Jb + (size_t(i) * cap_b + k) * 8, 8); }
    pb += b.ids_b.size();
    if (Js) { double* d = Js + size_t(i) * 16; std::memset(d, 0, 16 * sizeof(double));
      copy_block(jac[pb], d, 8); copy_block(jac[pb + 1], d + 8, 6); copy_block(jac[pb + 2], d + 14, 2); }
    if (Jrho) copy_block(jac[pb + 3], Jrho + 2 * size_t(i), 2);
  }
  auto toc = std::chrono::steady_clock::now();
  if (eval_seconds) *eval_seconds = std::chrono::duration<double>(toc - tic).count();
  int worst = KTO_OK;
  for (int i = 0; i < n; ++i) { if (status) status[i] = st[i]; if (st[i] != KTO_OK) worst = st[i]; }
  if (worst != KTO_OK && !first_err.empty()) g_last_error = first_err;
  return worst;
}

// LiftingRsCameraMeasurement residual blocks (lifting_rscamera_measurement.h:151-229).  vt[n]: the current value of each measurement's
// frame-normalised row time (its own parameter block, bounds [0, 1]); vt_orig = obs_uv.y / rows (:68).
//   r[3n]; ids_a/Ja (3 x size_a blocks), ids_b/Jb (3 x 4), Jvt[3n], Jrho[3n], Js[24n] (may be NULL); i0_* as kto_static_rs_residuals (observation at the lifted time).
int kto_lifting_rs_residuals(const kto_traj* tr, const kto_sensor* cam, const kto_camera* cmeta, int n, const double* obs_uv,
                             const double* obs_t0, const double* ref_uv, const double* ref_t0, const int* lm_idx, const double* rho,
                             const double* vt, const double* weight, int jac_mode, int nthreads, double* r, int cap_a, int* ids_a,
                             double* Ja, int cap_b, int* ids_b, double* Jb, double* Jvt, double* Jrho, int* i0_ref_a, int* i0_obs_a,
                             int* i0_ref_b, int* i0_obs_b, int* status, double* eval_seconds, double* Js) {
  TrajData td(*tr);
  CameraMeta cm; cm.readout = cmeta->readout; cm.rows = cmeta->rows; cm.cols = cmeta->cols;
  for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) cm.K[a][c] = cmeta->K[3 * a + c];
  cm.model = cmeta->model; cm.wc[0] = cmeta->wc[0]; cm.wc[1] = cmeta->wc[1]; cm.gamma = cmeta->gamma;
  std::vector<Block> blocks(n);
  std::vector<int> st(n, KTO_OK);
  std::string first_err;
  for (int i = 0; i < n; ++i) {
    st[i] = guarded([&] {
      double t1, t2;                                                                     // :162-177
      if (ref_t0[i] <= obs_t0[i]) { t1 = ref_t0[i]; t2 = obs_t0[i]; } else { t1 = obs_t0[i]; t2 = ref_t0[i]; }
      if (!cam->d_locked) { t1 -= cam->max_time_offset; t2 += cam->max_time_offset; }
      const double margin = 1e-3;
      add_trajectory(td, {{t1 - margin, t1 + cm.readout + margin}, {t2 - margin, t2 + cm.readout + margin}}, blocks[i]);
      add_sensor(*cam, blocks[i]);
      blocks[i].params.push_back(vt + i); blocks[i].sizes.push_back(1); blocks[i].constant.push_back(0);              // :200-205
      blocks[i].params.push_back(rho + lm_idx[i]); blocks[i].sizes.push_back(1); blocks[i].constant.push_back(0);     // :209-215
      if (int(blocks[i].ids_a.size()) > cap_a && ids_a) throw std::length_error("cap_a too small");
      if (int(blocks[i].ids_b.size()) > cap_b && ids_b) throw std::length_error("cap_b too small");
    });
    if (st[i] != KTO_OK && first_err.empty()) first_err = g_last_error;
  }
  const int sa = td.has_a() ? td.size_a() : 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  auto tic = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; ++i) {
    if (st[i] != KTO_OK) continue;
    const Block& b = blocks[i];
    LiftingRsFunctor f{&b, cm, weight ? weight[i] : 1.0, ref_uv + 2 * i, obs_uv + 2 * i, ref_t0[i], obs_t0[i], obs_uv[2 * i + 1] / double(cm.rows)};
    std::vector<std::vector<double>> jac;
    st[i] = guarded([&] { evaluate_block(f, b, 3, jac_mode, r + 3 * i, jac); });
    if (st[i] != KTO_OK) continue;
    const double row_delta = cm.readout / double(cm.rows);
    const double t_ref = ref_t0[i] + cam->time_offset + ref_uv[2 * i + 1] * row_delta;
    const double t_obs = obs_t0[i] + cam->time_offset + vt[i] * cm.readout;
    if (i0_ref_a) i0_ref_a[i] = td.has_a() ? locate_knot(b.meta.a, b.ids_a, t_ref) : -1;
    if (i0_obs_a) i0_obs_a[i] = td.has_a() ? locate_knot(b.meta.a, b.ids_a, t_obs) : -1;
    if (i0_ref_b) i0_ref_b[i] = td.has_b() ? locate_knot(b.meta.b, b.ids_b, t_ref) : -1;
    if (i0_obs_b) i0_obs_b[i] = td.has_b() ? locate_knot(b.meta.b, b.ids_b, t_obs) : -1;
    if (ids_a) for (int k = 0; k < cap_a; ++k) ids_a[size_t(i) * cap_a + k] = k < int(b.ids_a.size()) ? b.ids_a[k] : -1;
    if (ids_b) for (int k = 0; k < cap_b; ++k) ids_b[size_t(i) * cap_b + k] = k < int(b.ids_b.size()) ? b.ids_b[k] : -1;
    if (jac_mode == 0) continue;
    size_t pb = 0;
    if (Ja) { std::memset(Ja + size_t(i) * cap_a * 3 * sa, 0, sizeof(double) * cap_a * 3 * sa);
      for (size_t k = 0; k < b.ids_a.size(); ++k) copy_block(jac[pb + k], Ja + (size_t(i) * cap_a + k) * 3 * sa, 3 * sa); }
    pb += b.ids_a.size();
    if (Jb) { std::memset(Jb + size_t(i) * cap_b * 12, 0, sizeof(double) * cap_b * 12);
      for (size_t k = 0; k < b.ids_b.size(); ++k) copy_block(jac[pb + k], Jb + (size_t(i) * cap_b + k) * 12, 12); }
    pb += b.ids_b.size();
    // sensor blocks (sensors.h:135-165): q_ct (3 x 4) | p_ct (3 x 3) | time offset (3 x 1)
    if (Js) { double* d = Js + size_t(i) * 24; std::memset(d, 0, 24 * sizeof(double));
      copy_block(jac[pb], d, 12); copy_block(jac[pb + 1], d + 12, 9); copy_block(jac[pb + 2], d + 21, 3); }
    if (Jvt) copy_block(jac[pb + 3], Jvt + 3 * size_t(i), 3);
    if (Jrho) copy_block(jac[pb + 4], Jrho + 3 * size_t(i), 3);
  }
  auto toc = std::chrono::steady_clock::now();
  if (eval_seconds) *eval_seconds = std::chrono::duration<double>(toc - tic).count();
  int worst = KTO_OK;
  for (int i = 0; i < n; ++i) { if (status) status[i] = st[i]; if (st[i] != KTO_OK) worst = st[i]; }
  if (worst != KTO_OK && !first_err.empty()) g_last_error = first_err;
  return worst;
}

static CameraMeta camera_meta_of(const kto_camera* cmeta) {
  CameraMeta cm; cm.readout = cmeta->readout; cm.rows = cmeta->rows; cm.cols = cmeta->cols;
  for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) cm.K[a][c] = cmeta->K[3 * a + c];
  cm.model = cmeta->model; cm.wc[0] = cmeta->wc[0]; cm.wc[1] = cmeta->wc[1]; cm.gamma = cmeta->gamma;
  return cm;
}
// CameraView::EvaluateProjection(X, dX, true) / Unproject on doubles (sensors/pinhole_camera.h:47-67, atan_camera.h:54-103):
// what python/tests/test_cameras.py:32-75 exercises.
void kto_camera_project(const kto_camera* cmeta, const double* X, const double* dX, double* y, double* dy) {
  const CameraMeta cm = camera_meta_of(cmeta);
  camera_project<double>(cm, Vec3<double>{X[0], X[1], X[2]}, Vec3<double>{dX[0], dX[1], dX[2]}, true, y, dy);
}
void kto_camera_unproject(const kto_camera* cmeta, const double* y, double* X) {
  const CameraMeta cm = camera_meta_of(cmeta);
  const Vec3<double> v = camera_unproject<double>(cm, y);
  X[0] = v.x; X[1] = v.y; X[2] = v.z;
}

// ceres::HuberLoss(a) + ceres::internal::Corrector (un-vendored Ceres 1.x), applied by Ceres
// to the residual block AFTER Evaluate: r (nres) and J (nres x ncols row-major) are corrected in place; returns rho(s).
double kto_huber_correct(double a, int nres, int ncols, double* r, double* J) {
  double s = 0; for (int i = 0; i < nres; ++i) s += r[i] * r[i];
  const double b = a * a; double rho[3];
  if (s > b) { const double rr = std::sqrt(s); rho[0] = 2.0 * a * rr - b; rho[1] = std::max(std::numeric_limits<double>::min(), a / rr); rho[2] = -rho[1] / (2.0 * s); }
  else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  const double sqrt_rho1 = std::sqrt(rho[1]);
  double residual_scaling, alpha_sq_norm;
  if ((s == 0.0) || (rho[2] <= 0.0)) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; }
  else { const double D = 1.0 + 2.0 * s * rho[2] / rho[1]; const double alpha = 1.0 - std::sqrt(D); residual_scaling = sqrt_rho1 / (1 - alpha); alpha_sq_norm = alpha / s; }
  if (J) {
    if (alpha_sq_norm == 0.0) { for (int i = 0; i < nres * ncols; ++i) J[i] *= sqrt_rho1; }
    else for (int c = 0; c < ncols; ++c) {
      double rtj = 0; for (int i = 0; i < nres; ++i) rtj += J[i * ncols + c] * r[i];
      for (int i = 0; i < nres; ++i) J[i * ncols + c] = sqrt_rho1 * (J[i * ncols + c] - alpha_sq_norm * r[i] * rtj);
    }
  }
  for (int i = 0; i < nres; ++i) r[i] *= residual_scaling;
  return rho[0];
}

// Sophus::LocalParameterizationSE3::Plus (uniform_se3_spline_trajectory.h:25-33): T * exp(delta), delta = [upsilon; omega]
void kto_se3_plus(const double* T_raw, const double* delta, double* out) {
  SE3<double> T{{T_raw[0], T_raw[1], T_raw[2], T_raw[3]}, {T_raw[4], T_raw[5], T_raw[6]}};
  Vec6<double> d; for (int i = 0; i < 6; ++i) d.d[i] = delta[i];
  SE3<double> R = se3_mul(T, se3_exp(d));
  out[0] = R.q.x; out[1] = R.q.y; out[2] = R.q.z; out[3] = R.q.w; out[4] = R.t.x; out[5] = R.t.y; out[6] = R.t.z;
}

// Structure only: knot ids per ordered span list (spline_base.h:361-404), for tests of the segment rule.
int kto_spline_structure(double dt, double t0, int nspans, const double* spans, int cap, int* ids, int* nids, double* seg_t0, int* seg_n, int* nseg) {
  return guarded([&] {
    std::vector<std::pair<double, double>> times; for (int i = 0; i < nspans; ++i) times.push_back({spans[2 * i], spans[2 * i + 1]});
    SplineMeta meta; std::vector<int> k;
    spline_add_to_problem(dt, t0, times, meta, k);
    if (int(k.size()) > cap) throw std::length_error("cap too small");
    for (size_t i = 0; i < k.size(); ++i) ids[i] = k[i];
    *nids = int(k.size()); *nseg = int(meta.segments.size());
    for (size_t s = 0; s < meta.segments.size(); ++s) { seg_t0[s] = meta.segments[s].t0; seg_n[s] = int(meta.segments[s].n); }
  });
}

}  // extern "C"
