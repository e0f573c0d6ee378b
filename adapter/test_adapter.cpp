// test_adapter.cpp -- drives D2GpuSolver the way D2Estimator drives its SolverWrapper (d2estimator.cpp:604-685): the
// reference's own factor objects + residual descriptors go in through addResidual(), solve(fn) runs with a properties
// callback, results come back through the raw state pointers.  Input: a synthetic window dumped by tests/test_adapter.py.
//   test_adapter <in.bin> <out.bin> marshal   -- no CUDA call: dumps the flat C-ABI records the adapter built
//   test_adapter <in.bin> <out.bin> solve     -- full solve on the GPU, dumps the solved state + report
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "d2gpu_solver.hpp"
#include "d2vins_params.hpp"
#include "factors/imu_factor.h"
#include "factors/projectionOneFrameTwoCamFactor.h"
#include "factors/projectionTwoFrameOneCamDepthFactor.h"
#include "factors/projectionTwoFrameOneCamFactor.h"
#include "factors/projectionTwoFrameTwoCamFactor.h"

// statics whose home translation units need ROS (d2common/src/d2imu.cpp:8-9, d2vins/src/d2vins_params.cpp)
namespace D2Common {
Vector3d IMUData::Gravity = Vector3d(0., 0., 9.805);
Eigen::Matrix<double, 18, 18> IntegrationBase::noise = Eigen::Matrix<double, 18, 18>::Zero();
}  // namespace D2Common
namespace D2VINS { D2VINSConfig *params = nullptr; }

using namespace D2VINS;
using namespace D2Common;

namespace {
struct Reader {
  FILE *f;
  template <class T> void get(T *p, size_t n) { if (n && fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }
  template <class T> std::vector<T> vec(size_t n) { std::vector<T> v(n); get(v.data(), n); return v; }
  int64_t i64() { int64_t v; get(&v, 1); return v; }
  double f64() { double v; get(&v, 1); return v; }
};
struct Writer {
  FILE *f;
  template <class T> void put(const T *p, size_t n) { if (n) fwrite(p, sizeof(T), n, f); }
  void i64(int64_t v) { put(&v, 1); }
  void f64(double v) { put(&v, 1); }
};

class StubState : public D2EstimatorState {
 public:
  StubState() : D2EstimatorState(0) {}
  static StatePtr mk(const double *v, int n) { StatePtr p = makeSharedStateArray(n); memcpy(p.get(), v, sizeof(double) * n); return p; }
  void addPose(FrameIdType id, const double *v) { _frame_pose_state[id] = mk(v, 7); }
  void addExt(int id, const double *v) { _camera_extrinsic_state[id] = mk(v, 7); }
  void addSb(FrameIdType id, const double *v) { spd_bias_state[id] = mk(v, 9); }
  void addLm(LandmarkIdType id, double v) { landmark_state[id] = mk(&v, 1); }
  void addTd(int drone, double v) { td_state[drone] = mk(&v, 1); }
};

// stand-in for D2VINS::PriorFactor (prior_factor.cpp needs the whole estimator tree): same interface, r = e0 + J dx with
// the block-wise dx of prior_factor.cpp:57-68
class StubPrior : public PriorFactor {
  std::vector<ParamInfo> keep; int m; std::vector<double> J, e0;
 public:
  StubPrior(std::vector<ParamInfo> k, int mm, std::vector<double> Jm, std::vector<double> e) : keep(std::move(k)), m(mm), J(std::move(Jm)), e0(std::move(e)) {
    set_num_residuals(m);
    for (auto &p : keep) mutable_parameter_block_sizes()->push_back(p.size);
  }
  std::vector<ParamInfo> getKeepParams() const override { return keep; }
  int getEffParamsDim() const override { return m; }
  bool Evaluate(double const *const *x, double *r, double **jac) const override {
    std::vector<double> dx(m, 0.0);
    int off = 0;
    for (size_t k = 0; k < keep.size(); k++) {
      const double *x0 = keep[k].data_copied.data();
      if (IsSE3(keep[k].type)) {
        for (int q = 0; q < 3; q++) dx[off + q] = x[k][q] - x0[q];
        Eigen::Quaterniond q0(x0[6], x0[3], x0[4], x0[5]), q(x[k][6], x[k][3], x[k][4], x[k][5]);
        Eigen::Quaterniond e = Utility::positify(q0.inverse() * q);
        dx[off + 3] = 2 * e.x(); dx[off + 4] = 2 * e.y(); dx[off + 5] = 2 * e.z();
      } else for (int q = 0; q < keep[k].size; q++) dx[off + q] = x[k][q] - x0[q];
      off += keep[k].eff_size;
    }
    for (int i = 0; i < m; i++) { double s = e0[i]; for (int j = 0; j < m; j++) s += J[(size_t)i * m + j] * dx[j]; r[i] = s; }
    if (jac) {
      off = 0;
      for (size_t k = 0; k < keep.size(); k++) {
        if (jac[k]) for (int i = 0; i < m; i++) for (int c = 0; c < keep[k].size; c++) jac[k][(size_t)i * keep[k].size + c] = c < keep[k].eff_size ? J[(size_t)i * m + off + c] : 0.0;
        off += keep[k].eff_size;
      }
    }
    return true;
  }
};
Eigen::Vector3d v3(const double *p) { return Eigen::Vector3d(p[0], p[1], p[2]); }
}  // namespace

int main(int argc, char **argv) {
  if (argc < 4) { fprintf(stderr, "usage: test_adapter in.bin out.bin marshal|solve\n"); return 1; }
  Reader R{fopen(argv[1], "rb")};
  if (!R.f) { perror(argv[1]); return 1; }
  const bool do_solve = strcmp(argv[3], "solve") == 0;
  const int64_t np = R.i64(), ne = R.i64(), nsb = R.i64(), nl = R.i64(), nobs = R.i64(), nimu = R.i64(), pm = R.i64(), pnb = R.i64(), iters = R.i64();
  auto frame_ids = R.vec<int64_t>(np); auto poses = R.vec<double>(np * 7); auto pose_const = R.vec<uint8_t>(np);
  auto cam_ids = R.vec<int64_t>(ne); auto exts = R.vec<double>(ne * 7); auto ext_const = R.vec<uint8_t>(ne);
  auto sb_ids = R.vec<int64_t>(nsb); auto sbs = R.vec<double>(nsb * 9);
  const double td = R.f64(); const int64_t td_const = R.i64();
  auto lm_ids = R.vec<int64_t>(nl); auto lms = R.vec<double>(nl);
  auto obs = R.vec<d2ba_proj_obs>(nobs); auto imu = R.vec<d2ba_imu>(nimu);
  auto pJ = R.vec<double>(pm * pm); auto pe0 = R.vec<double>(pm); auto prefs = R.vec<d2ba_blockref>(pnb); auto px0 = R.vec<double>(R.i64());
  fclose(R.f);

  // d2vins_params.cpp:172-180
  const double focal = 460.0, depth_sqrt_inf = 20.0;
  ProjectionTwoFrameOneCamFactor::sqrt_info = focal / 1.5 * Matrix2d::Identity();
  ProjectionOneFrameTwoCamFactor::sqrt_info = focal / 1.5 * Matrix2d::Identity();
  ProjectionTwoFrameTwoCamFactor::sqrt_info = focal / 1.5 * Matrix2d::Identity();
  ProjectionTwoFrameOneCamDepthFactor::sqrt_info = focal / 1.5 * Matrix3d::Identity();
  ProjectionTwoFrameOneCamDepthFactor::sqrt_info(2, 2) = depth_sqrt_inf;

  StubState state;
  for (int64_t i = 0; i < np; i++) state.addPose(frame_ids[i], &poses[i * 7]);
  for (int64_t i = 0; i < ne; i++) state.addExt((int)cam_ids[i], &exts[i * 7]);
  for (int64_t i = 0; i < nsb; i++) state.addSb(sb_ids[i], &sbs[i * 9]);
  for (int64_t i = 0; i < nl; i++) state.addLm(lm_ids[i], lms[i]);
  state.addTd(0, td);

  D2GpuSolverConfig cfg; cfg.max_num_iterations = (int)iters; cfg.max_solver_time = 0.0; cfg.focal_length = focal; cfg.depth_sqrt_inf = depth_sqrt_inf;
  D2GpuSolver solver(&state, cfg);
  solver.reset();
  auto loss = std::make_shared<ceres::HuberLoss>(1.0);   // d2estimator.cpp:764
  // setupImuFactors (d2estimator.cpp:700-736)
  for (auto &m : imu) {
    auto pre = std::make_shared<IntegrationBase>(Vector3d(0, 0, 0), Vector3d(0, 0, 0), v3(m.linearized_ba), v3(m.linearized_bg));
    pre->sum_dt = m.sum_dt; pre->delta_p = v3(m.delta_p); pre->delta_v = v3(m.delta_v);
    pre->delta_q = Eigen::Quaterniond(m.delta_q[3], m.delta_q[0], m.delta_q[1], m.delta_q[2]);
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) { pre->jacobian(i, j) = m.jacobian[i * 15 + j]; pre->covariance(i, j) = m.covariance[i * 15 + j]; }
    auto info = std::make_shared<ImuResInfo>();
    info->frame_ida = m.frame_a; info->frame_idb = m.frame_b; info->cost_function = std::make_shared<IMUFactor>(pre);
    solver.addResidual(info);
  }
  // setupLandmarkFactors (d2estimator.cpp:796-874): one reference factor object per residual block
  for (auto &o : obs) {
    switch (o.type) {
      case D2BA_PROJ_2F1C: {
        auto info = std::make_shared<LandmarkTwoFrameOneCamResInfo>();
        info->frame_ida = o.frame_a; info->frame_idb = o.frame_b; info->landmark_id = o.landmark_id; info->camera_id = o.cam_a;
        info->cost_function = std::make_shared<ProjectionTwoFrameOneCamFactor>(v3(o.pts_i), v3(o.pts_j), v3(o.vel_i), v3(o.vel_j), o.td_i, o.td_j);
        info->loss_function = loss; solver.addResidual(info); break;
      }
      case D2BA_PROJ_2F1C_DEPTH: {
        auto info = std::make_shared<LandmarkTwoFrameOneCamResInfo>();
        info->frame_ida = o.frame_a; info->frame_idb = o.frame_b; info->landmark_id = o.landmark_id; info->camera_id = o.cam_a; info->enable_depth_mea = true;
        info->cost_function = std::make_shared<ProjectionTwoFrameOneCamDepthFactor>(v3(o.pts_i), v3(o.pts_j), v3(o.vel_i), v3(o.vel_j), o.td_i, o.td_j, o.depth);
        info->loss_function = loss; solver.addResidual(info); break;
      }
      case D2BA_PROJ_2F2C: {
        auto info = std::make_shared<LandmarkTwoFrameTwoCamResInfo>();
        info->frame_ida = o.frame_a; info->frame_idb = o.frame_b; info->landmark_id = o.landmark_id; info->camera_id_a = o.cam_a; info->camera_id_b = o.cam_b;
        info->cost_function = std::make_shared<ProjectionTwoFrameTwoCamFactor>(v3(o.pts_i), v3(o.pts_j), v3(o.vel_i), v3(o.vel_j), o.td_i, o.td_j);
        info->loss_function = loss; solver.addResidual(info); break;
      }
      case D2BA_PROJ_1F2C: {
        auto info = std::make_shared<LandmarkOneFrameTwoCamResInfo>();
        info->frame_ida = o.frame_a; info->landmark_id = o.landmark_id; info->camera_id_a = o.cam_a; info->camera_id_b = o.cam_b;
        info->cost_function = std::make_shared<ProjectionOneFrameTwoCamFactor>(v3(o.pts_i), v3(o.pts_j), v3(o.vel_i), v3(o.vel_j), o.td_i, o.td_j);
        info->loss_function = loss; solver.addResidual(info); break;
      }
      default: fprintf(stderr, "obs type %d not used by this test\n", o.type); return 3;
    }
  }
  // setupPriorFactor (d2estimator.cpp:888-897)
  if (pm > 0) {
    std::vector<ParamInfo> keep; size_t xo = 0;
    for (auto &r : prefs) {
      ParamInfo pi;
      if (r.kind == D2BA_POSE) pi = createFramePose(&state, r.id);
      else if (r.kind == D2BA_SPEED_BIAS) pi = createSpeedBias(&state, r.id);
      else if (r.kind == D2BA_EXTRINSIC) pi = createExtrinsic(&state, (int)r.id);
      else { fprintf(stderr, "prior kind\n"); return 3; }
      pi.data_copied = Eigen::Map<const VectorXd>(&px0[xo], pi.size); xo += pi.size;
      keep.push_back(pi);
    }
    solver.addResidual(std::make_shared<PriorResInfo>(std::make_shared<StubPrior>(keep, (int)pm, pJ, pe0)));
  }
  // the properties callback: D2Estimator::setStateProperties' rules on the bookkeeping problem
  StatePropertyInputs sp;
  bool any_free_ext = false;
  for (auto c : ext_const) if (!c) any_free_ext = true;
  sp.estimate_extrinsic = any_free_ext; sp.not_estimate_first_extrinsic = any_free_ext && ext_const[0];
  sp.estimate_td = td_const == 0; sp.has_prior = pm > 0; sp.always_fixed_first_pose = pm > 0 && pose_const[0];
  for (int64_t i = 0; i < ne; i++) sp.extrinsics_of_self_in_order.push_back(state.getExtrinsicState((int)cam_ids[i]).get());
  sp.td = state.getTdState(0).get(); sp.first_pose_of_self = state.getPoseState(frame_ids[0]).get();
  auto fn = [&]() { applyStateProperties(solver.getProblem(), sp); };

  Writer W{fopen(argv[2], "wb")};
  if (!do_solve) {
    D2GpuMarshalled m; std::string err;
    if (!solver.marshal(fn, m, err)) { fprintf(stderr, "marshal failed: %s\n", err.c_str()); return 4; }
    W.i64(m.pose_ids.size()); W.put(m.pose_ids.data(), m.pose_ids.size()); W.put(m.poses.data(), m.poses.size()); W.put(m.pose_const.data(), m.pose_const.size());
    W.i64(m.ext_ids.size()); W.put(m.ext_ids.data(), m.ext_ids.size()); W.put(m.exts.data(), m.exts.size()); W.put(m.ext_const.data(), m.ext_const.size());
    W.i64(m.sb_ids.size()); W.put(m.sb_ids.data(), m.sb_ids.size()); W.put(m.sbs.data(), m.sbs.size());
    W.i64(m.lm_ids.size()); W.put(m.lm_ids.data(), m.lm_ids.size()); W.put(m.lms.data(), m.lms.size());
    W.f64(m.td); W.i64(m.td_const);
    W.i64(m.obs.size()); W.put(m.obs.data(), m.obs.size()); W.i64(m.imu.size()); W.put(m.imu.data(), m.imu.size());
    W.i64(m.prior_m); W.put(m.prior_J.data(), m.prior_J.size()); W.put(m.prior_e0.data(), m.prior_e0.size());
  } else {
    SolverReport rep = solver.solve(fn);
    if (!rep.succ) { fprintf(stderr, "solve failed: %s\n", rep.message.c_str()); return 5; }
    W.i64(rep.total_iterations); W.f64(rep.initial_cost); W.f64(rep.final_cost); W.f64(rep.state_changes);
    for (int64_t i = 0; i < np; i++) W.put(state.getPoseState(frame_ids[i]).get(), 7);
    for (int64_t i = 0; i < nsb; i++) W.put(state.getSpdBiasState(sb_ids[i]).get(), 9);
    for (int64_t i = 0; i < nl; i++) W.put(state.getLandmarkState(lm_ids[i]).get(), 1);
  }
  fclose(W.f);
  return 0;
}
