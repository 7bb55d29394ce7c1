// Minimal stand-in for <ceres/ceres.h> -- TEST INFRASTRUCTURE ONLY.  Ceres is not installed in this image (nor is the reference buildable), so
// integration/batched_backend.h is compile- and run-checked against the two interfaces it derives from, declared here with the signatures of
// Ceres 1.14 / 2.x (ceres/evaluation_callback.h, ceres/cost_function.h).  Nothing else of Ceres is declared: the binding uses nothing else.
#pragma once
#include <cstdint>
#include <vector>

namespace ceres {

class EvaluationCallback {
 public:
  virtual ~EvaluationCallback() {}
  virtual void PrepareForEvaluation(bool evaluate_jacobians, bool new_evaluation_point) = 0;
};

class CostFunction {
 public:
  CostFunction() : num_residuals_(0) {}
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int32_t>& parameter_block_sizes() const { return parameter_block_sizes_; }
  int num_residuals() const { return num_residuals_; }

 protected:
  std::vector<int32_t>* mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
  void set_num_residuals(int n) { num_residuals_ = n; }

 private:
  std::vector<int32_t> parameter_block_sizes_;
  int num_residuals_;
};

}  // namespace ceres
