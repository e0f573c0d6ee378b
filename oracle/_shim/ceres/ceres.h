// Minimal stand-in for the ceres-solver declarations the D2SLAM factor sources derive from (TEST INFRASTRUCTURE,
// oracle/_ref build only).  Only interfaces: CostFunction / SizedCostFunction / LocalParameterization / LossFunction.
// ceres::Solve itself is NOT provided -- the minimizer stays the restated ("ASSUMED") part of the oracle.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <Eigen/Dense>
namespace ceres {
typedef int int32;
class CostFunction {
 public:
  CostFunction() : num_residuals_(0) {}
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const = 0;
  const std::vector<int32> &parameter_block_sizes() const { return parameter_block_sizes_; }
  int num_residuals() const { return num_residuals_; }
 protected:
  std::vector<int32> *mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
  void set_num_residuals(int n) { num_residuals_ = n; }
 private:
  std::vector<int32> parameter_block_sizes_;
  int num_residuals_;
};
template <int kNumResiduals, int... Ns>
class SizedCostFunction : public CostFunction {
 public:
  SizedCostFunction() { set_num_residuals(kNumResiduals); *mutable_parameter_block_sizes() = std::vector<int32>{Ns...}; }
  virtual ~SizedCostFunction() {}
};
class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual bool Plus(const double *x, const double *delta, double *x_plus_delta) const = 0;
  virtual bool ComputeJacobian(const double *x, double *jacobian) const = 0;
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
};
class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};
// rho(s) of HuberLoss as documented by Ceres (loss_function.h): s <= a^2 ? s : 2 a sqrt(s) - a^2
class HuberLoss : public LossFunction {
 public:
  explicit HuberLoss(double a) : a_(a), b_(a * a) {}
  void Evaluate(double s, double rho[3]) const override {
    if (s > b_) { const double r = std::sqrt(s); rho[0] = 2.0 * a_ * r - b_; rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r); rho[2] = -rho[1] / (2.0 * s); }
    else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  }
 private:
  const double a_, b_;
};
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TrustRegionStrategyType { LEVENBERG_MARQUARDT, DOGLEG };
enum Ownership { DO_NOT_TAKE_OWNERSHIP, TAKE_OWNERSHIP };
typedef LocalParameterization Manifold;
// Parameter-block bookkeeping of ceres::Problem (no residual storage, no solve): what D2Estimator::setStateProperties
// pokes through SolverWrapper::getProblem() -- HasParameterBlock / SetParameterBlockConstant / SetParameterization -- and
// what an adapter reads back with IsParameterBlockConstant.  Same member signatures as ceres 2.1.
class Problem {
 public:
  struct Options {
    Ownership cost_function_ownership = TAKE_OWNERSHIP, loss_function_ownership = TAKE_OWNERSHIP,
              local_parameterization_ownership = TAKE_OWNERSHIP, manifold_ownership = TAKE_OWNERSHIP;
  };
  Problem() {}
  explicit Problem(const Options &) {}
  void AddParameterBlock(double *values, int size) { blocks_[values].size = size; }
  void AddParameterBlock(double *values, int size, LocalParameterization *lp) { blocks_[values].size = size; blocks_[values].lp = lp; }
  bool HasParameterBlock(const double *values) const { return blocks_.count(const_cast<double *>(values)) != 0; }
  void SetParameterBlockConstant(const double *values) { blocks_.at(const_cast<double *>(values)).constant = true; }
  void SetParameterBlockVariable(double *values) { blocks_.at(values).constant = false; }
  bool IsParameterBlockConstant(const double *values) const { return blocks_.at(const_cast<double *>(values)).constant; }
  void SetParameterization(double *values, LocalParameterization *lp) { blocks_.at(values).lp = lp; }
  void SetManifold(double *values, Manifold *m) { blocks_.at(values).lp = m; }
  const LocalParameterization *GetParameterization(const double *values) const { return blocks_.at(const_cast<double *>(values)).lp; }
  int ParameterBlockSize(const double *values) const { return blocks_.at(const_cast<double *>(values)).size; }
  int NumParameterBlocks() const { return (int)blocks_.size(); }
  void *AddResidualBlock(CostFunction *, LossFunction *, const std::vector<double *> &ps) { for (double *p : ps) blocks_[p]; return nullptr; }
 private:
  struct Blk { int size = 0; bool constant = false; LocalParameterization *lp = nullptr; };
  std::map<double *, Blk> blocks_;
};
struct Solver {
  struct Options {
    LinearSolverType linear_solver_type = DENSE_SCHUR;
    TrustRegionStrategyType trust_region_strategy_type = DOGLEG;
    int num_threads = 1; int max_num_iterations = 50; double max_solver_time_in_seconds = 1e9;
    bool minimizer_progress_to_stdout = false;
  };
  struct Summary { int num_successful_steps = 0, num_unsuccessful_steps = 0; double initial_cost = 0, final_cost = 0, total_time_in_seconds = 0; std::string BriefReport() const { return ""; } std::string FullReport() const { return ""; } };
};
using std::cos; using std::sin; using std::floor; using std::sqrt; using std::atan2; using std::abs;
}  // namespace ceres
