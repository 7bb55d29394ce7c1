// integration/batched_backend.h -- the reference-side binding of libkontiki_b200.so, complete and compilable (INTEGRATION.md section 1 quotes it).
//
// What a maintainer of hovren/kontiki would add to cpplib/include/kontiki/: a ceres::EvaluationCallback that evaluates EVERY measurement of the
// estimator in one ktk_evaluate per parameter point, and thin ceres::CostFunction objects that copy their row out of that batch.  It replaces
//   * ceres::DynamicAutoDiffCostFunction<Residual> created in GyroscopeMeasurement::AddToEstimator (measurements/gyroscope_measurement.h:75-105),
//     AccelerometerMeasurement::AddToEstimator (accelerometer_measurement.h:77-108), StaticRsCameraMeasurement::AddToEstimator
//     (static_rscamera_measurement.h:130-198), and
//   * the per-block Evaluate loop of ceres::Solve (trajectory_estimator.h:38-64),
// with the same parameter blocks in the same order.  Nothing in this repository's product depends on this header; tests/test_cpp_binding.py compiles
// it against integration/ceres_stub (Ceres is not installed here) and runs integration/example_main.cc on the GPU against the Python binding.
#pragma once
#include <ceres/ceres.h>

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "kontiki_b200.h"

namespace kontiki {

// Owns the ktk_problem, the packed parameter point and the host-side result buffers of one estimator.
class BatchedBackend : public ceres::EvaluationCallback {
 public:
  explicit BatchedBackend(int device) { check(ktk_problem_create(device, &p_)); }
  ~BatchedBackend() override { if (p_) ktk_problem_destroy(p_); }
  BatchedBackend(const BatchedBackend&) = delete;
  BatchedBackend& operator=(const BatchedBackend&) = delete;

  // UniformSE3SplineTrajectory: knots are separate `new double[7]` blocks (entity/paramstore/dynamic_pstore.h:27); remember their addresses so
  // that the point can be packed before every evaluation.  compat_zero_dB = 1 is what the reference's accelerometer Jet path computes.
  void SetSpline(double dt, double t0, const std::vector<double*>& knot_blocks, int compat_zero_dB = 1) {
    knots_ = knot_blocks;
    packed_.resize(7 * knots_.size());
    check(ktk_set_se3_spline(p_, dt, t0, (int32_t)knots_.size(), compat_zero_dB));
  }

  // Rows are appended while the measurements' AddToEstimator runs; Finalize() creates one group per measurement type.  Each returns the row index.
  int AddGyroscopeRow(double t, const double w[3], double weight) { return imu_[0].add(t, w, weight); }
  int AddAccelerometerRow(double t, const double a[3], double weight) { return imu_[1].add(t, a, weight); }
  int AddStaticRsRow(const double obs_uv[2], double obs_t0, const double ref_uv[2], double ref_t0, double* rho_block, double weight, double huber_c) {
    int lm = -1;      // Landmark::inverse_depth_ptr() (sfm/landmark_impl.h:60-62): one block per landmark, shared by its observations
    for (size_t l = 0; l < rho_blocks_.size() && lm < 0; ++l) if (rho_blocks_[l] == rho_block) lm = (int)l;
    if (lm < 0) { lm = (int)rho_blocks_.size(); rho_blocks_.push_back(rho_block); }
    cam_.obs_uv.insert(cam_.obs_uv.end(), obs_uv, obs_uv + 2); cam_.obs_t0.push_back(obs_t0);
    cam_.ref_uv.insert(cam_.ref_uv.end(), ref_uv, ref_uv + 2); cam_.ref_t0.push_back(ref_t0);
    cam_.lm.push_back(lm); cam_.w.push_back(weight); cam_.huber.push_back(huber_c);
    return (int)cam_.w.size() - 1;
  }
  void SetImu(const ktk_sensor& imu) { imu_sensor_ = imu; }
  void SetCamera(const ktk_camera& cam) { camera_ = cam; }

  // After the last AddToEstimator: create the groups and the result buffers.
  void Finalize() {
    outs_.clear(); group_of_[0] = group_of_[1] = group_of_[2] = -1;
    for (int which = 0; which < 2; ++which) {
      ImuRows& m = imu_[which];
      if (m.t.empty()) continue;
      const int g = which == 0 ? ktk_add_gyroscope(p_, &imu_sensor_, (int64_t)m.t.size(), m.t.data(), m.y.data(), m.w.data())
                               : ktk_add_accelerometer(p_, &imu_sensor_, (int64_t)m.t.size(), m.t.data(), m.y.data(), m.w.data());
      check_group(g); group_of_[which] = g;
      m.r.resize(3 * m.t.size()); m.J.resize(84 * m.t.size()); m.i0.resize(m.t.size());
      ktk_group_out o{}; o.r = m.r.data(); o.J = m.J.data(); o.i0 = m.i0.data();
      outs_.push_back(o);
    }
    if (!cam_.w.empty()) {
      const int64_t n = (int64_t)cam_.w.size();
      const int g = ktk_add_static_rs(p_, &camera_, n, cam_.obs_uv.data(), cam_.obs_t0.data(), cam_.ref_uv.data(), cam_.ref_t0.data(), cam_.lm.data(),
                                      cam_.w.data(), cam_.huber.data());
      check_group(g); group_of_[2] = g;
      cam_.r.resize(2 * n); cam_.J.resize(114 * n); cam_.i0_ref.resize(n); cam_.i0_obs.resize(n);
      ktk_group_out o{}; o.r = cam_.r.data(); o.J = cam_.J.data(); o.i0 = cam_.i0_ref.data(); o.i0_b = cam_.i0_obs.data();
      outs_.push_back(o);
    }
    rho_.resize(rho_blocks_.size());
  }

  // ceres::EvaluationCallback: called once per parameter point before the residual-block loop.
  void PrepareForEvaluation(bool evaluate_jacobians, bool new_evaluation_point) override {
    if (!new_evaluation_point && (!evaluate_jacobians || have_jacobians_)) return;
    for (size_t k = 0; k < knots_.size(); ++k) std::copy(knots_[k], knots_[k] + 7, packed_.data() + 7 * k);
    for (size_t l = 0; l < rho_blocks_.size(); ++l) rho_[l] = *rho_blocks_[l];
    const uint32_t flags = KTK_EVAL_RESIDUALS | (evaluate_jacobians ? KTK_EVAL_JACOBIANS : 0u);      // the Huber loss stays with Ceres (LossFunction)
    check(ktk_evaluate(p_, packed_.data(), rho_.empty() ? nullptr : rho_.data(), (int64_t)rho_.size(), flags, outs_.data()));
    have_jacobians_ = evaluate_jacobians;
  }

  const ktk_group_out& imu_out(int which) const { return outs_[group_of_[which]]; }
  const ktk_group_out& camera_out() const { return outs_[group_of_[2]]; }
  int camera_group() const { return group_of_[2]; }
  const ktk_problem* problem() const { return p_; }

 private:
  struct ImuRows {
    std::vector<double> t, y, w, r, J; std::vector<int32_t> i0;
    int add(double t_, const double y_[3], double w_) { t.push_back(t_); y.insert(y.end(), y_, y_ + 3); w.push_back(w_); return (int)t.size() - 1; }
  };
  struct CamRows { std::vector<double> obs_uv, obs_t0, ref_uv, ref_t0, w, huber, r, J; std::vector<int32_t> lm, i0_ref, i0_obs; };
  // std::range_error where the reference throws it from inside Evaluate (spline_base.h:196-201): pybind11 maps it to ValueError as today
  static void check(int code) {
    if (code == KTK_OK) return;
    const std::string msg = ktk_last_error();
    if (code == KTK_ERANGE) throw std::range_error(msg);
    throw std::runtime_error(msg);
  }
  static void check_group(int g) { if (g < 0) check(g); }

  ktk_problem* p_ = nullptr;
  std::vector<double*> knots_, rho_blocks_;
  std::vector<double> packed_, rho_;
  ktk_sensor imu_sensor_{{0, 0, 0, 1}, {0, 0, 0}, 0.0, 0.1, 1, 1, 1};
  ktk_camera camera_{};
  ImuRows imu_[2];
  CamRows cam_;
  std::vector<ktk_group_out> outs_;
  int group_of_[3] = {-1, -1, -1};
  bool have_jacobians_ = false;
};

// Replaces ceres::DynamicAutoDiffCostFunction<Residual> at gyroscope_measurement.h:79 / accelerometer_measurement.h:81: same parameter blocks
// (the 4 knots of the segment, spline_base.h:371-403, then the sensor's q_ct, p_ct, time_offset, sensors.h:139-161), same residual count; Evaluate only copies.
class BatchedImuCost : public ceres::CostFunction {
 public:
  BatchedImuCost(const BatchedBackend* b, int which, int row) : b_(b), which_(which), row_(row) {
    set_num_residuals(3);
    for (int k = 0; k < 4; ++k) mutable_parameter_block_sizes()->push_back(7);
    for (int s : {4, 3, 1}) mutable_parameter_block_sizes()->push_back(s);
  }
  bool Evaluate(double const* const*, double* residuals, double** jacobians) const override {
    const ktk_group_out& o = b_->imu_out(which_);
    std::copy(o.r + 3 * row_, o.r + 3 * row_ + 3, residuals);
    if (jacobians) {
      for (int k = 0; k < 4; ++k)
        if (jacobians[k]) std::copy(o.J + 84 * row_ + 21 * k, o.J + 84 * row_ + 21 * (k + 1), jacobians[k]);      // row-major 3 x 7, Ceres' layout
      const int sizes[3] = {12, 9, 3};
      for (int s = 0; s < 3; ++s) if (jacobians[4 + s]) std::fill(jacobians[4 + s], jacobians[4 + s] + sizes[s], 0.0);      // locked sensor blocks
    }
    return true;
  }
  int first_knot() const { return b_->imu_out(which_).i0[row_]; }      // the 4 knot blocks are knots first_knot() .. + 3

 private:
  const BatchedBackend* b_; int which_, row_;
};

// Replaces the cost function at static_rscamera_measurement.h:134.  Parameter blocks: the STRUCTURAL knot list of the residual (all knots of its one or
// two segments, :137-168; ktk_get_structure returns it), the camera's q_ct, p_ct, time_offset, then the landmark's inverse depth.
class BatchedStaticRsCost : public ceres::CostFunction {
 public:
  BatchedStaticRsCost(const BatchedBackend* b, int row, const std::vector<int32_t>& knot_ids) : b_(b), row_(row), ids_(knot_ids) {
    set_num_residuals(2);
    for (size_t k = 0; k < ids_.size(); ++k) mutable_parameter_block_sizes()->push_back(7);
    for (int s : {4, 3, 1, 1}) mutable_parameter_block_sizes()->push_back(s);
  }
  bool Evaluate(double const* const*, double* residuals, double** jacobians) const override {
    const ktk_group_out& o = b_->camera_out();
    std::copy(o.r + 2 * row_, o.r + 2 * row_ + 2, residuals);
    if (!jacobians) return true;
    const double* J = o.J + 114 * (size_t)row_;      // [ref 4 x (2x7) | obs 4 x (2x7) | d r/d rho (2)]
    const size_t nk = ids_.size();
    for (size_t k = 0; k < nk; ++k) if (jacobians[k]) std::fill(jacobians[k], jacobians[k] + 14, 0.0);
    for (int w = 0; w < 2; ++w) {
      const int base = w == 0 ? o.i0[row_] : o.i0_b[row_];
      for (int k = 0; k < 4; ++k)
        for (size_t c = 0; c < nk; ++c)
          if (ids_[c] == base + k && jacobians[c]) for (int e = 0; e < 14; ++e) jacobians[c][e] += J[56 * w + 14 * k + e];      // windows may overlap: sum
    }
    const int sizes[3] = {8, 6, 2};
    for (int s = 0; s < 3; ++s) if (jacobians[nk + s]) std::fill(jacobians[nk + s], jacobians[nk + s] + sizes[s], 0.0);           // locked camera blocks
    if (jacobians[nk + 3]) { jacobians[nk + 3][0] = J[112]; jacobians[nk + 3][1] = J[113]; }
    return true;
  }

 private:
  const BatchedBackend* b_; int row_; std::vector<int32_t> ids_;
};

}  // namespace kontiki
