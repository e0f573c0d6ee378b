// ref_driver.cpp -- C entry points over the REFERENCE's own factor classes (oracle/_ref/libd2ref.so).
//
// TEST INFRASTRUCTURE ONLY.  The reference sources are compiled unmodified from /root/reference by oracle/Makefile.ref:
//   d2vins/src/factors/projectionTwoFrameOneCamFactor.cpp, projectionTwoFrameTwoCamFactor.cpp,
//   projectionOneFrameTwoCamFactor.cpp, projectionTwoFrameOneCamDepthFactor.cpp, imu_factor.h (+ d2common
//   integration_base.h, utils.hpp), d2common/src/solver/consenus_factor.cpp, pose_local_parameterization.cpp.
// Their third-party dependencies (Eigen, ceres, ROS, swarm_msgs, OpenCV, spdlog) are absent from this image and are
// stood in for by oracle/_shim (interfaces + a small eager matrix library).  This file only constructs the reference's
// objects from flat arrays and calls their Evaluate / Plus / propagate -- it contains no factor arithmetic of its own.
// It pins oracle/orc_factors.c (tests/test_ref_pin.py) and generates tests/golden/ref_factors.npz.
#include <cstring>
#include <memory>

#include <d2common/integration_base.h>
#include <d2common/solver/consenus_factor.h>
#include <d2common/solver/pose_local_parameterization.h>
#include <d2common/solver/RelPoseFactor.hpp>
#include <d2common/solver/BaseParamResInfo.hpp>
#include <d2common/solver/ConsensusSolver.hpp>
#include <pthread.h>
#include <thread>
#include <d2common/utils.hpp>

#include "d2vins_params.hpp"
#include "factors/imu_factor.h"
#include "factors/prior_factor.h"
#include "estimator/marginalization/marginalization.hpp"
#include "../include/d2ba.h"
#include "posegraph_g2o.hpp"
#include <random>
#include "factors/projectionOneFrameTwoCamFactor.h"
#include "factors/projectionTwoFrameOneCamDepthFactor.h"
#include "factors/projectionTwoFrameOneCamFactor.h"
#include "factors/projectionTwoFrameTwoCamFactor.h"

// statics whose home translation units (d2common/src/d2imu.cpp, d2vins/src/d2vins_params.cpp) need ROS / OpenCV file
// storage: defined here with the values those files assign (d2imu.cpp:8-9, d2vins_params.cpp:58-74,172-180)
namespace D2Common {
Vector3d IMUData::Gravity = Vector3d(0., 0., 9.805);
Eigen::Matrix<double, 18, 18> IntegrationBase::noise = Eigen::Matrix<double, 18, 18>::Zero();
}  // namespace D2Common
namespace D2VINS { D2VINSConfig *params = nullptr; }

using namespace D2VINS;
using namespace D2Common;

extern "C" {

// d2vins_params.cpp:58-74 and :172-180
void ref_configure(double focal_length, double depth_sqrt_inf, double g_norm, double acc_n, double gyr_n, double acc_w, double gyr_w) {
  Eigen::Matrix<double, 18, 18> noise = Eigen::Matrix<double, 18, 18>::Zero();
  noise.block<3, 3>(0, 0) = (acc_n * acc_n) * Eigen::Matrix3d::Identity();
  noise.block<3, 3>(3, 3) = (gyr_n * gyr_n) * Eigen::Matrix3d::Identity();
  noise.block<3, 3>(6, 6) = (acc_n * acc_n) * Eigen::Matrix3d::Identity();
  noise.block<3, 3>(9, 9) = (gyr_n * gyr_n) * Eigen::Matrix3d::Identity();
  noise.block<3, 3>(12, 12) = (acc_w * acc_w) * Eigen::Matrix3d::Identity();
  noise.block<3, 3>(15, 15) = (gyr_w * gyr_w) * Eigen::Matrix3d::Identity();
  IntegrationBase::noise = noise;
  IMUData::Gravity = Vector3d(0., 0., g_norm);
  ProjectionTwoFrameOneCamFactor::sqrt_info = focal_length / 1.5 * Matrix2d::Identity();
  ProjectionOneFrameTwoCamFactor::sqrt_info = focal_length / 1.5 * Matrix2d::Identity();
  ProjectionTwoFrameTwoCamFactor::sqrt_info = focal_length / 1.5 * Matrix2d::Identity();
  ProjectionTwoFrameOneCamDepthFactor::sqrt_info = focal_length / 1.5 * Matrix3d::Identity();
  ProjectionTwoFrameOneCamDepthFactor::sqrt_info(2, 2) = depth_sqrt_inf;
}

static Vector3d v3(const double *p) { return Vector3d(p[0], p[1], p[2]); }

// type: d2ba_proj_type (0 2F1C, 1 2F2C, 2 1F2C, 3 2F1C_DEPTH).  params in the factor's own block order, J[k] row-major
// (rows x block size) or NULL.  tangent_base_out (6) returns the constructor's tangent base.  Returns the residual size.
int ref_proj_eval(int type, const double *pts_i, const double *pts_j, const double *vel_i, const double *vel_j, double td_i, double td_j,
                  double depth_j, const double *const *params, double *residuals, double **jacobians, double *tangent_base_out) {
  std::unique_ptr<ceres::CostFunction> f;
  const double *tb = nullptr;
  switch (type) {
    case 0: { auto *p = new ProjectionTwoFrameOneCamFactor(v3(pts_i), v3(pts_j), v3(vel_i), v3(vel_j), td_i, td_j); tb = nullptr; f.reset(p);
              if (tangent_base_out) for (int r = 0; r < 2; r++) for (int c = 0; c < 3; c++) tangent_base_out[r * 3 + c] = p->tangent_base(r, c); break; }
    case 1: { auto *p = new ProjectionTwoFrameTwoCamFactor(v3(pts_i), v3(pts_j), v3(vel_i), v3(vel_j), td_i, td_j); f.reset(p);
              if (tangent_base_out) for (int r = 0; r < 2; r++) for (int c = 0; c < 3; c++) tangent_base_out[r * 3 + c] = p->tangent_base(r, c); break; }
    case 2: { auto *p = new ProjectionOneFrameTwoCamFactor(v3(pts_i), v3(pts_j), v3(vel_i), v3(vel_j), td_i, td_j); f.reset(p);
              if (tangent_base_out) for (int r = 0; r < 2; r++) for (int c = 0; c < 3; c++) tangent_base_out[r * 3 + c] = p->tangent_base(r, c); break; }
    case 3: { auto *p = new ProjectionTwoFrameOneCamDepthFactor(v3(pts_i), v3(pts_j), v3(vel_i), v3(vel_j), td_i, td_j, depth_j); f.reset(p);
              if (tangent_base_out) for (int r = 0; r < 2; r++) for (int c = 0; c < 3; c++) tangent_base_out[r * 3 + c] = p->tangent_base(r, c); break; }
    default: return -1;
  }
  (void)tb;
  if (!f->Evaluate(params, residuals, jacobians)) return -2;
  return f->num_residuals();
}

// IntegrationBase midpoint pre-integration (integration_base.h:95-199): acc/gyr are (n+1) x 3 with row 0 = acc_0 / gyr_0.
// out: sum_dt, delta_p(3), delta_q(4, xyzw), delta_v(3), jacobian(225 row-major), covariance(225 row-major)
static IntegrationBasePtr make_preint(int n, const double *dt, const double *acc, const double *gyr, const double *ba, const double *bg) {
  IntegrationBasePtr pre = std::make_shared<IntegrationBase>(v3(acc), v3(gyr), v3(ba), v3(bg));
  for (int k = 0; k < n; k++) pre->push_back(dt[k], v3(acc + 3 * (k + 1)), v3(gyr + 3 * (k + 1)));
  return pre;
}
void ref_preintegrate(int n, const double *dt, const double *acc, const double *gyr, const double *ba, const double *bg, double *out) {
  IntegrationBasePtr pre = make_preint(n, dt, acc, gyr, ba, bg);
  out[0] = pre->sum_dt;
  for (int k = 0; k < 3; k++) { out[1 + k] = pre->delta_p(k); out[8 + k] = pre->delta_v(k); }
  out[4] = pre->delta_q.x(); out[5] = pre->delta_q.y(); out[6] = pre->delta_q.z(); out[7] = pre->delta_q.w();
  for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) { out[11 + i * 15 + j] = pre->jacobian(i, j); out[11 + 225 + i * 15 + j] = pre->covariance(i, j); }
}

// IMUFactor built from explicit pre-integration results (the C ABI's d2ba_imu record); Evaluate with params
// {pose_i(7), sb_i(9), pose_j(7), sb_j(9)}; J[k] row-major 15 x {7,9,7,9}; sqrt_info_out 225 row-major (imu_factor.h:29)
int ref_imu_eval(double sum_dt, const double *delta_p, const double *delta_q_xyzw, const double *delta_v, const double *lin_ba, const double *lin_bg,
                 const double *jacobian, const double *covariance, const double *const *params, double *residuals, double **jacobians,
                 double *sqrt_info_out) {
  IntegrationBasePtr pre = std::make_shared<IntegrationBase>(Vector3d(0, 0, 0), Vector3d(0, 0, 0), v3(lin_ba), v3(lin_bg));
  pre->sum_dt = sum_dt; pre->delta_p = v3(delta_p); pre->delta_v = v3(delta_v);
  pre->delta_q = Eigen::Quaterniond(delta_q_xyzw[3], delta_q_xyzw[0], delta_q_xyzw[1], delta_q_xyzw[2]);
  for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) { pre->jacobian(i, j) = jacobian[i * 15 + j]; pre->covariance(i, j) = covariance[i * 15 + j]; }
  IMUFactor f(pre);
  if (sqrt_info_out) {
    // the factor keeps sqrt_info private: recompute it with the very expression of its constructor (imu_factor.h:29)
    Eigen::Matrix<double, 15, 15> si = Eigen::LLT<Eigen::Matrix<double, 15, 15>>(pre->covariance.inverse()).matrixL().transpose();
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) sqrt_info_out[i * 15 + j] = si(i, j);
  }
  return f.Evaluate(params, residuals, jacobians) ? 15 : -2;
}

// ConsenusPoseFactor (consenus_factor.cpp:6-51): residual 6, J 6 x 7 row-major
int ref_consensus_eval(const double *t_ref, const double *q_ref_xyzw, const double *t_tilde, const double *theta_tilde, double rho_T, double rho_theta,
                       const double *pose, double *r6, double *J6x7) {
  ConsenusPoseFactor f(v3(t_ref), Eigen::Quaterniond(q_ref_xyzw[3], q_ref_xyzw[0], q_ref_xyzw[1], q_ref_xyzw[2]), v3(t_tilde), v3(theta_tilde), rho_T, rho_theta);
  const double *params[1] = {pose};
  double *jac[1] = {J6x7};
  return f.Evaluate(params, r6, J6x7 ? jac : nullptr) ? 6 : -2;
}

// RelPoseFactorAD (d2common/include/d2common/solver/RelPoseFactor.hpp:68-135), the pose-graph factor of d2pgo: the
// reference's own templated functor, evaluated with T = double (residual) and with T = Jet (oracle/_shim/ceres: dual
// numbers) -- i.e. the Jacobians are the exact derivatives of the reference's residual code w.r.t. the 7 ambient pose
// parameters [p, q(xyzw)], what ceres autodiff hands to the manifold.  J: 6 x 7 row-major per pose.
int ref_relpose_ad_eval(const double *pose_a, const double *pose_b, const double *rel7, const double *sqrt_info36, double *r6, double *Ja6x7, double *Jb6x7) {
  Eigen::Matrix6d S;
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) S(i, j) = sqrt_info36[i * 6 + j];
  D2Common::RelPoseFactorAD f(Swarm::Pose(rel7), S);
  const double *params[2] = {pose_a, pose_b};
  double *jac[2] = {Ja6x7, Jb6x7};
  return ceres::AutoDiffEvaluate<6, 7, 7>(f, params, r6, Ja6x7 ? jac : nullptr) ? 6 : -2;
}

// Loss corrector: ResidualInfo::Evaluate(param_infos) (d2common/src/solver/BaseParamResInfo.cpp:46-92, compiled from the
// reference) applied to a cost function that returns the given residual / Jacobian, with ceres::HuberLoss(a) (the shim's
// restatement of the un-vendored ceres class).  One parameter block of n_par doubles; J row-major n_res x n_par.
namespace {
struct FixedCost : ceres::CostFunction {
  const double *r; const double *J; int nr, np;
  FixedCost(const double *r_, const double *J_, int nr_, int np_) : r(r_), J(J_), nr(nr_), np(np_) { set_num_residuals(nr_); mutable_parameter_block_sizes()->push_back(np_); }
  bool Evaluate(double const *const *, double *res, double **jac) const override {
    for (int i = 0; i < nr; i++) res[i] = r[i];
    if (jac && jac[0]) for (int i = 0; i < nr * np; i++) jac[0][i] = J[i];
    return true;
  }
};
struct PlainResidualInfo : D2Common::ResidualInfo {
  PlainResidualInfo() : D2Common::ResidualInfo(D2Common::NONE) {}
  bool relavant(const std::set<FrameIdType> &) const override { return false; }
  std::vector<D2Common::ParamInfo> paramsList(D2Common::D2State *) const override { return {}; }
};
}  // namespace
int ref_loss_correct(int n_res, int n_par, const double *r_in, const double *J_in, double huber_a, double *r_out, double *J_out) {
  PlainResidualInfo ri;
  ri.cost_function = std::make_shared<FixedCost>(r_in, J_in, n_res, n_par);
  if (huber_a > 0) ri.loss_function = std::make_shared<ceres::HuberLoss>(huber_a);
  D2Common::ParamInfo p;
  p.pointer = std::shared_ptr<double>(new double[n_par](), std::default_delete<double[]>());
  p.size = n_par; p.eff_size = n_par;
  ri.Evaluate(std::vector<D2Common::ParamInfo>{p}, false);
  for (int i = 0; i < n_res; i++) r_out[i] = ri.residuals(i);
  for (int i = 0; i < n_res; i++) for (int j = 0; j < n_par; j++) J_out[i * n_par + j] = ri.jacobians[0](i, j);
  return n_res;
}

// RelPoseFactor4D (RelPoseFactor.hpp:196-238), d2pgo's DEFAULT factor (pgo_pose_dof = PGO_POSE_4D, d2pgo_config.h): poses are
// [x y z yaw]; same evaluation scheme as ref_relpose_ad_eval (doubles / dual numbers).  J: 4 x 4 row-major per pose.
int ref_relpose4d_eval(const double *pose_a4, const double *pose_b4, const double *rel7, const double *sqrt_info16, double *r4, double *Ja4x4, double *Jb4x4) {
  Eigen::Matrix4d S;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) S(i, j) = sqrt_info16[i * 4 + j];
  D2Common::RelPoseFactor4D f(Swarm::Pose(rel7), S);
  const double *params[2] = {pose_a4, pose_b4};
  double *jac[2] = {Ja4x4, Jb4x4};
  return ceres::AutoDiffEvaluate<4, 4, 4>(f, params, r4, Ja4x4 ? jac : nullptr) ? 4 : -2;
}

// PoseLocalParameterization (pose_local_parameterization.cpp:13-38); its members are private virtuals of
// ceres::LocalParameterization, reached through the base interface exactly as ceres does
void ref_pose_plus(const double *x7, const double *delta6, double *out7) {
  PoseLocalParameterization p;
  const ceres::LocalParameterization &b = p;
  b.Plus(x7, delta6, out7);
}
void ref_pose_plus_jacobian(const double *x7, double *J7x6) {
  PoseLocalParameterization p;
  const ceres::LocalParameterization &b = p;
  b.ComputeJacobian(x7, J7x6);
}

// Utility helpers (d2common/include/d2common/utils.hpp:56-104, 213-228)
void ref_qleft_qright(const double *q_xyzw, double *L16, double *R16) {
  Eigen::Quaterniond q(q_xyzw[3], q_xyzw[0], q_xyzw[1], q_xyzw[2]);
  Eigen::Matrix4d L = Utility::Qleft(q), R = Utility::Qright(q);
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { L16[i * 4 + j] = L(i, j); R16[i * 4 + j] = R(i, j); }
}
void ref_average_quats(int n, const double *q_xyzw, double *out_xyzw) {
  std::vector<Eigen::Quaterniond> qs;
  for (int i = 0; i < n; i++) qs.emplace_back(q_xyzw[4 * i + 3], q_xyzw[4 * i], q_xyzw[4 * i + 1], q_xyzw[4 * i + 2]);
  Eigen::Quaterniond a = Utility::averageQuaterions(qs);
  out_xyzw[0] = a.x(); out_xyzw[1] = a.y(); out_xyzw[2] = a.z(); out_xyzw[3] = a.w();
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ ADMM bookkeeping replay
// The reference's OWN ConsensusSolver::solve loop (d2common/src/solver/ConsensusSolver.cpp:39-235, compiled unmodified:
// syncData -> updateGlobal, problem assembly, updateTilde, solveLocalStep) run for N agents on N threads, with
//   * broadcastData / waitForSync implemented as an in-process table + barrier (the reference's are LCM messages,
//     VINSConsenusSolver.cpp:11-120),
//   * ceres::Solve replaced by a hook that (a) snapshots param_global / param_tilde and evaluates the ConsenusPoseFactor
//     objects updateTilde just created, and (b) writes the NEXT prescribed local poses into the state -- i.e. a given
//     state trajectory is replayed through the reference's dual / global updates.
// Swarm::Pose::{averagePoses, DeltaPose, tangentSpace} are the ASSUMED stand-ins of oracle/_shim/swarm_msgs/Pose.h.
namespace Swarm {
Pose Pose::averagePoses(const std::vector<Pose> &poses) {
  Eigen::Vector3d p(0, 0, 0);
  std::vector<Eigen::Quaterniond> qs;
  for (const Pose &x : poses) { p += x.pos(); qs.push_back(x.att()); }
  p = p / (double)poses.size();
  return Pose(p, D2Common::Utility::averageQuaterions(qs));   // utils.hpp:213-228 (reference code)
}
}  // namespace Swarm
namespace {
struct ZeroCost : ceres::CostFunction {
  explicit ZeroCost(int nblk) { set_num_residuals(1); for (int i = 0; i < nblk; i++) mutable_parameter_block_sizes()->push_back(7); }
  bool Evaluate(double const *const *, double *res, double **) const override { res[0] = 0.0; return true; }
};
struct ReplayResInfo : D2Common::ResidualInfo {
  std::vector<D2Common::ParamInfo> infos;
  ReplayResInfo() : D2Common::ResidualInfo(D2Common::NONE) {}
  bool relavant(const std::set<FrameIdType> &) const override { return false; }
  std::vector<D2Common::ParamInfo> paramsList(D2Common::D2State *) const override { return infos; }
};
struct ReplayState : D2Common::D2State {
  ReplayState(int id, int n) : D2Common::D2State(id) { for (int a = 0; a < n; a++) all_drones.insert(a); }
};
struct ReplayShared {
  int n_agents, n_blocks, steps;
  const uint8_t *present; const double *traj;     // [a][b], [(steps+1)][a][b][7]
  double *z_out, *tilde_out, *res_out;            // [steps][a][b][7|6|6]
  std::vector<double> table;                      // [a][b][7] what every agent last broadcast
  pthread_barrier_t bar;
};
class ReplayAgent : public D2Common::ConsensusSolver {
  int me; ReplayShared *sh; std::vector<D2Common::StatePtr> ptr; int step = 0;
  void broadcastData() override {
    for (int b = 0; b < sh->n_blocks; b++) if (ptr[b]) memcpy(&sh->table[((size_t)me * sh->n_blocks + b) * 7], ptr[b].get(), 56);
  }
  void exchange() {
    pthread_barrier_wait(&sh->bar);               // every agent has broadcast
    for (int b = 0; b < sh->n_blocks; b++) {
      if (!ptr[b]) continue;
      for (int a = 0; a < sh->n_agents; a++) {
        if (!sh->present[a * sh->n_blocks + b]) continue;
        VectorXd v(7);
        for (int k = 0; k < 7; k++) v(k) = sh->table[((size_t)a * sh->n_blocks + b) * 7 + k];
        remote_params[ptr[b]][a] = v;
      }
    }
    pthread_barrier_wait(&sh->bar);               // every agent has read: the table may be overwritten
  }
  void waitForSync() override { exchange(); }
  void receiveAll() override { exchange(); }
  void on_solve(ceres::Problem *problem) {
    const int k = step++;
    for (int b = 0; b < sh->n_blocks; b++) {
      if (!ptr[b]) continue;
      const size_t o = ((size_t)k * sh->n_agents + me) * sh->n_blocks + b;
      const D2Common::ConsenusParamState &cs = consenus_params.at(ptr[b]);
      for (int q = 0; q < 7; q++) sh->z_out[o * 7 + q] = cs.param_global(q);
      for (int q = 0; q < 6; q++) sh->tilde_out[o * 6 + q] = cs.param_tilde(q);
    }
    for (const auto &rb : problem->residual_blocks()) {     // the factors updateTilde created for this step
      auto *f = dynamic_cast<ConsenusPoseFactor *>(rb.cost);
      if (!f) continue;
      for (int b = 0; b < sh->n_blocks; b++) {
        if (!ptr[b] || ptr[b].get() != rb.params[0]) continue;
        double r[6]; const double *pp[1] = {rb.params[0]};
        f->Evaluate(pp, r, nullptr);
        for (int q = 0; q < 6; q++) sh->res_out[(((size_t)k * sh->n_agents + me) * sh->n_blocks + b) * 6 + q] = r[q];
      }
    }
    for (int b = 0; b < sh->n_blocks; b++)        // "the local solve": the next prescribed poses
      if (ptr[b]) memcpy(ptr[b].get(), sh->traj + ((((size_t)(k + 1)) * sh->n_agents + me) * sh->n_blocks + b) * 7, 56);
  }
 public:
  ReplayAgent(D2Common::D2State *st, D2Common::ConsensusSolverConfig cfg, int me_, ReplayShared *sh_) : D2Common::ConsensusSolver(st, cfg, 0), me(me_), sh(sh_), ptr(sh_->n_blocks) {
    auto ri = std::make_shared<ReplayResInfo>();
    int nb = 0;
    for (int b = 0; b < sh->n_blocks; b++) {
      if (!sh->present[me * sh->n_blocks + b]) continue;
      ptr[b] = std::shared_ptr<double>(new double[7], std::default_delete<double[]>());
      memcpy(ptr[b].get(), sh->traj + (((size_t)me) * sh->n_blocks + b) * 7, 56);
      D2Common::ParamInfo p; p.pointer = ptr[b]; p.index = -1; p.size = 7; p.eff_size = 6; p.type = D2Common::POSE; p.id = b;
      p.data_copied = Map<VectorXd>(ptr[b].get(), 7);
      ri->infos.push_back(p); nb++;
    }
    ri->cost_function = std::make_shared<ZeroCost>(nb);
    addResidual(ri);
  }
  void run() {
    ceres::solve_hook() = [this](const ceres::Solver::Options &, ceres::Problem *p, ceres::Solver::Summary *) { on_solve(p); };
    solve();
    ceres::solve_hook() = nullptr;
  }
};
}  // namespace
extern "C" int ref_admm_replay(int n_agents, int n_blocks, int steps, double relaxation_alpha, double rho_T, double rho_theta, const uint8_t *present,
                               const double *traj, double *z_out, double *tilde_out, double *res_out) {
  ReplayShared sh;
  sh.n_agents = n_agents; sh.n_blocks = n_blocks; sh.steps = steps; sh.present = present; sh.traj = traj;
  sh.z_out = z_out; sh.tilde_out = tilde_out; sh.res_out = res_out; sh.table.assign((size_t)n_agents * n_blocks * 7, 0.0);
  pthread_barrier_init(&sh.bar, nullptr, n_agents);
  std::vector<std::unique_ptr<ReplayState>> states; std::vector<std::unique_ptr<ReplayAgent>> agents;
  for (int a = 0; a < n_agents; a++) {
    D2Common::ConsensusSolverConfig cfg;
    cfg.max_steps = steps; cfg.self_id = a; cfg.relaxation_alpha = relaxation_alpha; cfg.rho_frame_T = rho_T; cfg.rho_frame_theta = rho_theta; cfg.sync_for_averaging = true;
    states.emplace_back(new ReplayState(a, n_agents));
    agents.emplace_back(new ReplayAgent(states.back().get(), cfg, a, &sh));
  }
  std::vector<std::thread> th;
  for (int a = 0; a < n_agents; a++) th.emplace_back([&, a]() { agents[a]->run(); });
  for (auto &t : th) t.join();
  pthread_barrier_destroy(&sh.bar);
  return 0;
}

// ------------------------------------------------------------------------------------------------ marginalization prior
// PriorFactor (d2vins/src/factors/prior_factor.cpp:45-90 Evaluate, :132-177 toJacRes, compiled unmodified): constructed from
// the information form (A, b) and the kept blocks' linearisation points, evaluated at x.  kinds: 0 POSE, 1 EXTRINSIC,
// 2 SPEED_BIAS, 3 TD, 4 LANDMARK.  x0 / x: the blocks' values concatenated (7, 7, 9, 1, 1 doubles); J_out: per block a
// row-major m x size matrix, concatenated.
extern "C" int ref_prior_eval(int nblk, const int *kinds, const double *x0, const double *x, int m, const double *A, const double *b, double *r_out, double *J_out) {
  static D2VINSConfig cfg;                       // toJacRes reads params->debug_write_margin_matrix
  if (!D2VINS::params) D2VINS::params = &cfg;
  static const int kSize[5] = {7, 7, 9, 1, 1}, kEff[5] = {6, 6, 9, 1, 1};
  static const D2Common::ParamsType kType[5] = {D2Common::POSE, D2Common::EXTRINSIC, D2Common::SPEED_BIAS, D2Common::TD, D2Common::LANDMARK};
  std::vector<D2Common::ParamInfo> keep;
  std::vector<const double *> px;
  int off = 0, eff = 0;
  for (int i = 0; i < nblk; i++) {
    D2Common::ParamInfo p;
    const int sz = kSize[kinds[i]];
    p.pointer = std::shared_ptr<double>(new double[sz], std::default_delete<double[]>());
    memcpy(p.pointer.get(), x0 + off, sizeof(double) * sz);
    p.index = eff; p.size = sz; p.eff_size = kEff[kinds[i]]; p.type = kType[kinds[i]]; p.id = i;
    p.data_copied = Map<VectorXd>(p.pointer.get(), sz);
    keep.push_back(p); px.push_back(x + off);
    off += sz; eff += p.eff_size;
  }
  if (eff != m) return -1;
  MatrixXd Am(m, m); VectorXd bv(m);
  for (int i = 0; i < m; i++) { bv(i) = b[i]; for (int j = 0; j < m; j++) Am(i, j) = A[i * m + j]; }
  PriorFactor f(keep, Am, bv);
  std::vector<double *> jac;
  double *jp = J_out;
  for (int i = 0; i < nblk; i++) { jac.push_back(jp); jp += (size_t)m * kSize[kinds[i]]; }
  return f.Evaluate(px.data(), r_out, J_out ? jac.data() : nullptr) ? m : -2;
}

// ------------------------------------------------------------------------------------------------ marginalization
// The reference's OWN Marginalizer::marginalize (d2vins/src/estimator/marginalization/marginalization.cpp:13-286), its
// ResidualInfo classes (ParamResidualInfo.hpp / .cpp: relavant(), paramsList()), ResidualInfo::Evaluate with the Huber
// corrector, Utility::schurComplement and the PriorFactor constructor -- all compiled unmodified -- over the reference's
// factor objects of one window.  What is supplied here instead of d2vinsstate.cpp (the estimator's state class, which
// needs the whole front end): the five D2EstimatorState lookups those sources call, answered from a table.
namespace {
struct MargTable {
  std::map<FrameIdType, D2Common::StatePtr> pose, sb;
  std::map<int, D2Common::StatePtr> ext;
  std::map<LandmarkIdType, D2Common::StatePtr> lm;
  std::map<LandmarkIdType, FrameIdType> lm_base;
  D2Common::StatePtr td;
};
MargTable *g_marg = nullptr;
D2Common::StatePtr make_state(const double *v, int n) {
  D2Common::StatePtr p(new double[n], std::default_delete<double[]>());
  memcpy(p.get(), v, sizeof(double) * n);
  return p;
}
struct MargState : D2Common::D2State {      // D2State::getPoseState (d2state.hpp:87-95) reads _frame_pose_state
  explicit MargState(const MargTable &t) : D2Common::D2State(0) { for (auto &kv : t.pose) _frame_pose_state[kv.first] = kv.second; }
};
}  // namespace
namespace D2VINS {   // link-time stand-ins for d2vinsstate.cpp (`this` is never touched)
StatePtr D2EstimatorState::getExtrinsicState(int i) const { return g_marg->ext.at(i); }
StatePtr D2EstimatorState::getSpdBiasState(FrameIdType frame_id) const { return g_marg->sb.at(frame_id); }
StatePtr D2EstimatorState::getLandmarkState(LandmarkIdType landmark_id) const { return g_marg->lm.at(landmark_id); }
StatePtr D2EstimatorState::getTdState(int) { return g_marg->td; }
FrameIdType D2EstimatorState::getLandmarkBaseFrame(LandmarkIdType landmark_id) const { return g_marg->lm_base.at(landmark_id); }
}  // namespace D2VINS
extern "C" int ref_marginalize(int n_pose, const int64_t *pose_ids, const double *poses, const double *sb, int n_ext, const int64_t *cam_ids, const double *ext,
                               double td, int n_lm, const int64_t *lm_ids, const double *lm, const int64_t *lm_base, int n_obs, const d2ba_proj_obs *obs,
                               int n_imu, const d2ba_imu *imu, int prior_m, const double *prior_A, const double *prior_b, int prior_nblk,
                               const d2ba_blockref *prior_refs, const double *prior_x0, int n_remove, const int64_t *remove_ids, double huber_delta,
                               int *nblk_out, d2ba_blockref *refs_out, double *x0_out, int *m_out, double *J_out, double *e0_out, int max_m) {
  static D2VINSConfig cfg;
  D2VINS::params = &cfg;
  cfg.margin_enable_fej = false; cfg.margin_sparse_solver = true; cfg.remove_base_when_margin_remote = 2; cfg.verbose = false;
  cfg.enable_perf_output = false; cfg.debug_write_margin_matrix = false; cfg.landmark_param = D2VINSConfig::LM_INV_DEP;
  MargTable tab;
  for (int i = 0; i < n_pose; i++) { tab.pose[pose_ids[i]] = make_state(poses + 7 * i, 7); tab.sb[pose_ids[i]] = make_state(sb + 9 * i, 9); }
  for (int i = 0; i < n_ext; i++) tab.ext[(int)cam_ids[i]] = make_state(ext + 7 * i, 7);
  for (int i = 0; i < n_lm; i++) { tab.lm[lm_ids[i]] = make_state(lm + i, 1); tab.lm_base[lm_ids[i]] = lm_base[i]; }
  tab.td = make_state(&td, 1);
  g_marg = &tab;
  MargState state(tab);
  auto *est = reinterpret_cast<D2VINS::D2EstimatorState *>(static_cast<D2Common::D2State *>(&state));   // only ever used as a D2State / by the stand-ins above
  Marginalizer marg(est, nullptr);
  std::shared_ptr<ceres::LossFunction> loss;
  if (huber_delta > 0) loss = std::make_shared<ceres::HuberLoss>(huber_delta);
  for (int i = 0; i < n_obs; i++) {
    const d2ba_proj_obs &o = obs[i];
    switch (o.type) {
      case 0: marg.addResidualInfo(LandmarkTwoFrameOneCamResInfo::create(std::make_shared<ProjectionTwoFrameOneCamFactor>(v3(o.pts_i), v3(o.pts_j), v3(o.vel_i), v3(o.vel_j), o.td_i, o.td_j),
                                                                          loss, o.frame_a, o.frame_b, o.landmark_id, o.cam_a, false)); break;
      case 1: marg.addResidualInfo(LandmarkTwoFrameTwoCamResInfo::create(std::make_shared<ProjectionTwoFrameTwoCamFactor>(v3(o.pts_i), v3(o.pts_j), v3(o.vel_i), v3(o.vel_j), o.td_i, o.td_j),
                                                                          loss, o.frame_a, o.frame_b, o.landmark_id, o.cam_a, o.cam_b)); break;
      case 2: marg.addResidualInfo(LandmarkOneFrameTwoCamResInfo::create(std::make_shared<ProjectionOneFrameTwoCamFactor>(v3(o.pts_i), v3(o.pts_j), v3(o.vel_i), v3(o.vel_j), o.td_i, o.td_j),
                                                                          loss, o.frame_a, o.landmark_id, o.cam_a, o.cam_b)); break;
      case 3: marg.addResidualInfo(LandmarkTwoFrameOneCamResInfo::create(std::make_shared<ProjectionTwoFrameOneCamDepthFactor>(v3(o.pts_i), v3(o.pts_j), v3(o.vel_i), v3(o.vel_j), o.td_i, o.td_j, o.depth),
                                                                          loss, o.frame_a, o.frame_b, o.landmark_id, o.cam_a, true)); break;
      default: g_marg = nullptr; return -3;   // depth prior (autodiff) is not part of this replay
    }
  }
  for (int i = 0; i < n_imu; i++) {
    const d2ba_imu &m = imu[i];
    IntegrationBasePtr pre = std::make_shared<IntegrationBase>(Vector3d(0, 0, 0), Vector3d(0, 0, 0), v3(m.linearized_ba), v3(m.linearized_bg));
    pre->sum_dt = m.sum_dt; pre->delta_p = v3(m.delta_p); pre->delta_v = v3(m.delta_v);
    pre->delta_q = Eigen::Quaterniond(m.delta_q[3], m.delta_q[0], m.delta_q[1], m.delta_q[2]);
    for (int r = 0; r < 15; r++) for (int c = 0; c < 15; c++) { pre->jacobian(r, c) = m.jacobian[r * 15 + c]; pre->covariance(r, c) = m.covariance[r * 15 + c]; }
    marg.addResidualInfo(ImuResInfo::create(std::make_shared<IMUFactor>(pre), m.frame_a, m.frame_b));
  }
  static const int kSize[5] = {7, 7, 9, 1, 1}, kEff[5] = {6, 6, 9, 1, 1};
  static const D2Common::ParamsType kType[5] = {D2Common::POSE, D2Common::EXTRINSIC, D2Common::SPEED_BIAS, D2Common::TD, D2Common::LANDMARK};
  if (prior_m > 0) {
    std::vector<D2Common::ParamInfo> keep;
    int off = 0, eff = 0;
    for (int i = 0; i < prior_nblk; i++) {
      const int k = prior_refs[i].kind; const int64_t id = prior_refs[i].id;
      D2Common::ParamInfo p;
      p.pointer = k == 0 ? tab.pose.at(id) : k == 1 ? tab.ext.at((int)id) : k == 2 ? tab.sb.at(id) : k == 3 ? tab.td : tab.lm.at(id);
      p.index = eff; p.size = kSize[k]; p.eff_size = kEff[k]; p.type = kType[k]; p.id = id;
      p.data_copied = Map<const VectorXd>(prior_x0 + off, kSize[k]);
      keep.push_back(p); off += kSize[k]; eff += kEff[k];
    }
    if (eff != prior_m) { g_marg = nullptr; return -1; }
    MatrixXd A(prior_m, prior_m); VectorXd b(prior_m);
    for (int i = 0; i < prior_m; i++) { b(i) = prior_b[i]; for (int j = 0; j < prior_m; j++) A(i, j) = prior_A[i * prior_m + j]; }
    marg.addResidualInfo(std::make_shared<PriorResInfo>(std::make_shared<PriorFactor>(keep, A, b)));
  }
  std::set<FrameIdType> rem(remove_ids, remove_ids + n_remove);
  PriorFactorPtr prior = marg.marginalize(rem);
  g_marg = nullptr;
  if (!prior) return -2;
  std::vector<D2Common::ParamInfo> kp = prior->getKeepParams();
  const int m = prior->getEffParamsDim();
  if (m > max_m) return -4;
  *nblk_out = (int)kp.size(); *m_out = m;
  std::vector<const double *> px; std::vector<std::vector<double>> Jb(kp.size()); std::vector<double *> jp;
  int off = 0;
  for (size_t i = 0; i < kp.size(); i++) {
    int k = 0;
    for (int q = 0; q < 5; q++) if (kType[q] == kp[i].type) k = q;
    refs_out[i].kind = k; refs_out[i].pad = 0; refs_out[i].id = kp[i].id;
    for (int q = 0; q < kp[i].size; q++) x0_out[off + q] = kp[i].data_copied(q);
    px.push_back(x0_out + off); off += kp[i].size;
    Jb[i].assign((size_t)m * kp[i].size, 0.0); jp.push_back(Jb[i].data());
  }
  if (!prior->Evaluate(px.data(), e0_out, jp.data())) return -5;   // at the linearisation point: residual = e0, Jacobian blocks = columns of J
  for (size_t i = 0; i < kp.size(); i++)
    for (int r = 0; r < m; r++) for (int c = 0; c < kp[i].eff_size; c++) J_out[(size_t)r * m + kp[i].index + c] = Jb[i][(size_t)r * kp[i].size + c];
  return 0;
}

// ------------------------------------------------------------------------------------------------ g2o files (d2pgo)
// The reference's own reader / writer (d2pgo/test/posegraph_g2o.cpp:27-232, compiled unmodified): read_g2o_agent on one file,
// write_result_to_g2o.  (The three random-number externs it declares belong to d2pgo_test.cpp.)
namespace D2PGO { std::random_device rd; std::default_random_engine eng(0); std::normal_distribution<double> d(0, 1); }
extern "C" int ref_g2o_read(const char *path, int max_agent_id, int max_v, int max_e, int *nv, int *v_agent, int64_t *v_id, double *v_pose7,
                            int *ne, int *e_agent_a, int64_t *e_id_a, int *e_agent_b, int64_t *e_id_b, double *e_rel7, double *e_info36) {
  std::map<FrameIdType, D2BaseFramePtr> frames; std::vector<Swarm::LoopEdge> edges;
  D2PGO::read_g2o_agent(path, frames, edges, false, max_agent_id, -1, false);
  if ((int)frames.size() > max_v || (int)edges.size() > max_e) return -1;
  int i = 0;
  for (auto &kv : frames) { v_agent[i] = kv.second->drone_id; v_id[i] = kv.second->frame_id; kv.second->odom.pose().to_vector(v_pose7 + 7 * i); i++; }
  *nv = i; i = 0;
  for (auto &e : edges) {
    e_agent_a[i] = e.id_a; e_agent_b[i] = e.id_b; e_id_a[i] = e.keyframe_id_a; e_id_b[i] = e.keyframe_id_b; e.relative_pose.to_vector(e_rel7 + 7 * i);
    const Eigen::Matrix6d I6 = e.getInfoMat();
    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) e_info36[36 * i + r * 6 + c] = I6(r, c);
    i++;
  }
  *ne = i;
  return 0;
}
extern "C" int ref_g2o_write(const char *path, int nv, const int64_t *v_id, const double *v_pose7, int ne, const int64_t *e_id_a, const int64_t *e_id_b,
                             const double *e_rel7, const double *e_info36) {
  std::vector<D2BaseFramePtr> frames; std::vector<Swarm::LoopEdge> edges;
  for (int i = 0; i < nv; i++) { auto f = std::make_shared<D2BaseFrame>(); f->frame_id = v_id[i]; f->odom.pose() = Swarm::Pose(v_pose7 + 7 * i); frames.push_back(f); }
  for (int i = 0; i < ne; i++) {
    Eigen::Matrix6d I6;
    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) I6(r, c) = e_info36[36 * i + r * 6 + c];
    edges.emplace_back(e_id_a[i], e_id_b[i], Swarm::Pose(e_rel7 + 7 * i), I6);
  }
  D2PGO::write_result_to_g2o(path, frames, edges, false);
  return 0;
}
