// Minimal stand-in for the ceres-solver declarations the D2SLAM factor sources derive from (TEST INFRASTRUCTURE,
// oracle/_ref build only).  Only interfaces: CostFunction / SizedCostFunction / LocalParameterization / LossFunction.
// ceres::Solve itself is NOT provided -- the minimizer stays the restated ("ASSUMED") part of the oracle.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
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

// ---- forward-mode dual numbers + AutoDiffCostFunction (interface of ceres/jet.h, ceres/autodiff_cost_function.h): lets the
//      reference's templated functors (RelPoseFactorAD ...) run with T = Jet, i.e. their Jacobians are the exact derivatives
//      of the reference's own residual code.  Written here; not ceres code.
template <typename T, int N>
struct Jet {
  T a; T v[N];
  Jet() : a(T(0)) { for (int i = 0; i < N; i++) v[i] = T(0); }
  Jet(const T &x) : a(x) { for (int i = 0; i < N; i++) v[i] = T(0); }
  Jet(int x) : a(T(x)) { for (int i = 0; i < N; i++) v[i] = T(0); }
  Jet(const T &x, int k) : a(x) { for (int i = 0; i < N; i++) v[i] = T(i == k ? 1 : 0); }
  Jet &operator+=(const Jet &o) { a += o.a; for (int i = 0; i < N; i++) v[i] += o.v[i]; return *this; }
  Jet &operator-=(const Jet &o) { a -= o.a; for (int i = 0; i < N; i++) v[i] -= o.v[i]; return *this; }
  Jet &operator*=(const Jet &o) { *this = *this * o; return *this; }
  Jet &operator/=(const Jet &o) { *this = *this / o; return *this; }
};
template <typename T, int N> Jet<T, N> operator+(const Jet<T, N> &x) { return x; }
template <typename T, int N> Jet<T, N> operator-(const Jet<T, N> &x) { Jet<T, N> r; r.a = -x.a; for (int i = 0; i < N; i++) r.v[i] = -x.v[i]; return r; }
template <typename T, int N> Jet<T, N> operator+(const Jet<T, N> &x, const Jet<T, N> &y) { Jet<T, N> r; r.a = x.a + y.a; for (int i = 0; i < N; i++) r.v[i] = x.v[i] + y.v[i]; return r; }
template <typename T, int N> Jet<T, N> operator-(const Jet<T, N> &x, const Jet<T, N> &y) { Jet<T, N> r; r.a = x.a - y.a; for (int i = 0; i < N; i++) r.v[i] = x.v[i] - y.v[i]; return r; }
template <typename T, int N> Jet<T, N> operator*(const Jet<T, N> &x, const Jet<T, N> &y) { Jet<T, N> r; r.a = x.a * y.a; for (int i = 0; i < N; i++) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
template <typename T, int N> Jet<T, N> operator/(const Jet<T, N> &x, const Jet<T, N> &y) { Jet<T, N> r; const T inv = T(1) / y.a; r.a = x.a * inv; for (int i = 0; i < N; i++) r.v[i] = (x.v[i] - r.a * y.v[i]) * inv; return r; }
#define D2_JET_SCALAR_OPS(op) \
  template <typename T, int N> Jet<T, N> operator op(const Jet<T, N> &x, const T &s) { return x op Jet<T, N>(s); } \
  template <typename T, int N> Jet<T, N> operator op(const T &s, const Jet<T, N> &x) { return Jet<T, N>(s) op x; }
D2_JET_SCALAR_OPS(+) D2_JET_SCALAR_OPS(-) D2_JET_SCALAR_OPS(*) D2_JET_SCALAR_OPS(/)
#undef D2_JET_SCALAR_OPS
#define D2_JET_CMP(op) \
  template <typename T, int N> bool operator op(const Jet<T, N> &x, const Jet<T, N> &y) { return x.a op y.a; } \
  template <typename T, int N> bool operator op(const Jet<T, N> &x, const T &y) { return x.a op y; } \
  template <typename T, int N> bool operator op(const T &x, const Jet<T, N> &y) { return x op y.a; }
D2_JET_CMP(<) D2_JET_CMP(<=) D2_JET_CMP(>) D2_JET_CMP(>=) D2_JET_CMP(==) D2_JET_CMP(!=)
#undef D2_JET_CMP
template <typename T, int N> Jet<T, N> jet_chain(const Jet<T, N> &x, T f, T df) { Jet<T, N> r; r.a = f; for (int i = 0; i < N; i++) r.v[i] = df * x.v[i]; return r; }
template <typename T, int N> Jet<T, N> sqrt(const Jet<T, N> &x) { const T s = std::sqrt(x.a); return jet_chain(x, s, T(0.5) / s); }
template <typename T, int N> Jet<T, N> sin(const Jet<T, N> &x) { return jet_chain(x, std::sin(x.a), std::cos(x.a)); }
template <typename T, int N> Jet<T, N> cos(const Jet<T, N> &x) { return jet_chain(x, std::cos(x.a), -std::sin(x.a)); }
template <typename T, int N> Jet<T, N> atan2(const Jet<T, N> &y, const Jet<T, N> &x) {
  const T d = x.a * x.a + y.a * y.a; Jet<T, N> r; r.a = std::atan2(y.a, x.a);
  for (int i = 0; i < N; i++) r.v[i] = (x.a * y.v[i] - y.a * x.v[i]) / d;
  return r;
}
template <typename T, int N> Jet<T, N> asin(const Jet<T, N> &x) { return jet_chain(x, std::asin(x.a), T(1) / std::sqrt(T(1) - x.a * x.a)); }
template <typename T, int N> Jet<T, N> acos(const Jet<T, N> &x) { return jet_chain(x, std::acos(x.a), -T(1) / std::sqrt(T(1) - x.a * x.a)); }
template <typename T, int N> Jet<T, N> abs(const Jet<T, N> &x) { return x.a < T(0) ? -x : x; }
template <typename T, int N> Jet<T, N> fabs(const Jet<T, N> &x) { return x.a < T(0) ? -x : x; }
template <typename T, int N> bool isfinite(const Jet<T, N> &x) { return std::isfinite(x.a); }
template <typename T, int N> Jet<T, N> floor(const Jet<T, N> &x) { return Jet<T, N>(std::floor(x.a)); }   // piecewise constant: zero derivative

// Exact derivatives of a templated functor by dual numbers.  A free function on purpose: the class below does NOT
// instantiate the functor (RelPoseFactor.hpp's Create() helpers name AutoDiffCostFunction for factors -- 4-DoF, perturbation,
// 9-D rotation -- whose bodies need more of Eigen than oracle/_shim/Eigen provides); oracle/ref_driver.cpp calls this for the
// factor it pins.
template <int kNumResiduals, int... Ns, typename Functor>
bool AutoDiffEvaluate(const Functor &f, double const *const *parameters, double *residuals, double **jacobians) {
  constexpr int kBlocks = sizeof...(Ns);
  constexpr int kTotal = (Ns + ...);
  static_assert(kBlocks == 2, "stand-in: two parameter blocks");
  if (!jacobians) return f(parameters[0], parameters[1], residuals);
  typedef Jet<double, kTotal> J;
  const int sizes[kBlocks] = {Ns...};
  std::vector<J> x(kTotal); const J *ptr[kBlocks];
  for (int b = 0, o = 0; b < kBlocks; o += sizes[b], b++) { ptr[b] = x.data() + o; for (int k = 0; k < sizes[b]; k++) x[o + k] = J(parameters[b][k], o + k); }
  J r[kNumResiduals];
  if (!f(ptr[0], ptr[1], r)) return false;
  for (int i = 0; i < kNumResiduals; i++) residuals[i] = r[i].a;
  for (int b = 0, o = 0; b < kBlocks; o += sizes[b], b++)
    if (jacobians[b]) for (int i = 0; i < kNumResiduals; i++) for (int k = 0; k < sizes[b]; k++) jacobians[b][i * sizes[b] + k] = r[i].v[o + k];
  return true;
}
template <typename Functor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public SizedCostFunction<kNumResiduals, Ns...> {   // interface only (see AutoDiffEvaluate)
  std::unique_ptr<Functor> f_;
 public:
  explicit AutoDiffCostFunction(Functor *f) : f_(f) {}
  const Functor &functor() const { return *f_; }
  bool Evaluate(double const *const *, double *, double **) const override { return false; }
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
  struct ResBlk { CostFunction *cost; LossFunction *loss; std::vector<double *> params; };
  void *AddResidualBlock(CostFunction *c, LossFunction *l, const std::vector<double *> &ps) { for (double *p : ps) blocks_[p]; residual_blocks_.push_back(ResBlk{c, l, ps}); return nullptr; }
  void *AddResidualBlock(CostFunction *c, LossFunction *l, double *p0) { return AddResidualBlock(c, l, std::vector<double *>{p0}); }
  void *AddResidualBlock(CostFunction *c, LossFunction *l, double *p0, double *p1) { return AddResidualBlock(c, l, std::vector<double *>{p0, p1}); }
  const std::vector<ResBlk> &residual_blocks() const { return residual_blocks_; }   // stand-in only: what the caller assembled
 private:
  struct Blk { int size = 0; bool constant = false; LocalParameterization *lp = nullptr; };
  std::map<double *, Blk> blocks_;
  std::vector<ResBlk> residual_blocks_;
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

// ceres::NormalPrior(A, b): r = A (x - b), J = A (interface + the two lines of arithmetic of ceres/normal_prior.cc; stand-in)
class NormalPrior : public CostFunction {
 public:
  NormalPrior(const Eigen::MatrixXd &A, const Eigen::VectorXd &b) : A_(A), b_(b) { set_num_residuals(A_.rows()); mutable_parameter_block_sizes()->push_back(b_.rows()); }
  bool Evaluate(double const *const *p, double *r, double **J) const override {
    const int m = A_.rows(), n = A_.cols();
    for (int i = 0; i < m; i++) { double s = 0; for (int k = 0; k < n; k++) s += A_(i, k) * (p[0][k] - b_(k)); r[i] = s; }
    if (J && J[0]) for (int i = 0; i < m; i++) for (int k = 0; k < n; k++) J[0][i * n + k] = A_(i, k);
    return true;
  }
  const Eigen::MatrixXd &A() const { return A_; }
  const Eigen::VectorXd &b() const { return b_; }
 private:
  Eigen::MatrixXd A_; Eigen::VectorXd b_;
};

// ceres::Solve stand-in: there is no minimiser in this shim.  The call is forwarded to a hook the test driver installs
// (oracle/ref_driver.cpp replays a prescribed state trajectory through the reference's ADMM bookkeeping); without a hook
// it leaves the parameter blocks untouched.
typedef std::function<void(const Solver::Options &, Problem *, Solver::Summary *)> SolveHook;
inline SolveHook &solve_hook() { static thread_local SolveHook h; return h; }
inline void Solve(const Solver::Options &o, Problem *p, Solver::Summary *s) { if (solve_hook()) solve_hook()(o, p, s); }
}  // namespace ceres
