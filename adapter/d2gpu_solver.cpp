// d2gpu_solver.cpp -- see d2gpu_solver.hpp.  Host-side marshalling only: no numerics of the solve run here.
#include "d2gpu_solver.hpp"

#include <cmath>
#include <cstring>
#include <map>

// the reference's own factor classes (public members carry the per-observation constants)
#include "d2vins_params.hpp"
#include "factors/imu_factor.h"
#include "factors/projectionOneFrameTwoCamFactor.h"
#include "factors/projectionTwoFrameOneCamDepthFactor.h"
#include "factors/projectionTwoFrameOneCamFactor.h"
#include "factors/projectionTwoFrameTwoCamFactor.h"

namespace D2VINS {
using namespace D2Common;

D2GpuSolver::D2GpuSolver(D2State *_state, const D2GpuSolverConfig &cfg) : SolverWrapper(_state), cfg_(cfg) {}

D2GpuSolver::~D2GpuSolver() {
  if (h_) d2ba_destroy(h_);
}

void D2GpuSolver::reset() {
  SolverWrapper::reset();   // fresh bookkeeping problem, residual list cleared (SolverWrapper.cpp:20-24)
  if (h_) d2ba_reset(h_);
}

void D2GpuSolver::setConsensusSlots(const std::vector<d2ba_blockref> &refs, const std::vector<int32_t> &slots, int n_slots_global) {
  cons_refs_ = refs; cons_slots_ = slots; cons_n_slots_ = n_slots_global;
}

bool D2GpuSolver::ensureHandle() {
  if (h_) return true;
  d2ba_config c;
  d2ba_default_config(&c);
  c.device = cfg_.device; c.max_windows = 1; c.max_num_iterations = cfg_.max_num_iterations; c.consensus_max_steps = cfg_.consensus_max_steps;
  c.focal_length = cfg_.focal_length; c.depth_sqrt_inf = cfg_.depth_sqrt_inf; c.gravity_norm = cfg_.gravity_norm;
  c.rho_frame_T = cfg_.rho_frame_T; c.rho_frame_theta = cfg_.rho_frame_theta; c.rho_landmark = cfg_.rho_landmark; c.relaxation_alpha = cfg_.relaxation_alpha;
  c.max_solver_time_in_seconds = cfg_.max_solver_time;
  const int rc = d2ba_create(&c, &h_);
  if (rc) { err_ = "d2ba_create failed (a CUDA device is required; there is no CPU fallback), rc=" + std::to_string(rc); h_ = nullptr; return false; }
  return true;
}

namespace {
// distinct parameter blocks in first-seen order, keyed by the raw pointer ceres would have been given
struct BlockTable {
  std::map<double *, int> index;
  std::vector<int64_t> *ids; std::vector<double> *vals; std::vector<double *> *ptrs; int size;
  void add(double *p, int64_t id) {
    if (index.count(p)) return;
    index[p] = (int)ids->size();
    ids->push_back(id); ptrs->push_back(p);
    vals->insert(vals->end(), p, p + size);
  }
};

template <class F> void copy3(double *dst, const F &v) { dst[0] = v(0); dst[1] = v(1); dst[2] = v(2); }

template <class Factor> void fill_obs(d2ba_proj_obs &o, const Factor &f) {
  copy3(o.pts_i, f.pts_i); copy3(o.pts_j, f.pts_j); copy3(o.vel_i, f.velocity_i); copy3(o.vel_j, f.velocity_j);
  o.td_i = f.td_i; o.td_j = f.td_j;
}
}  // namespace

bool D2GpuSolver::marshal(std::function<void()> func_set_properties, D2GpuMarshalled &m, std::string &err) {
  m = D2GpuMarshalled();
  BlockTable poses{{}, &m.pose_ids, &m.poses, &m.pose_ptr, POSE_SIZE}, exts{{}, &m.ext_ids, &m.exts, &m.ext_ptr, POSE_SIZE},
      sbs{{}, &m.sb_ids, &m.sbs, &m.sb_ptr, FRAME_SPDBIAS_SIZE}, lms{{}, &m.lm_ids, &m.lms, &m.lm_ptr, INV_DEP_SIZE};
  // ---- 1. parameter blocks (what CeresSolver::solve hands to AddResidualBlock, SolverWrapper.cpp:27-33)
  for (auto &ri : residuals) {
    for (ParamInfo &pi : ri->paramsList(state)) {
      double *p = CheckGetPtr(pi.pointer);
      if (!problem->HasParameterBlock(p)) problem->AddParameterBlock(p, pi.size);
      switch (pi.type) {
        case POSE: poses.add(p, pi.id); break;
        case EXTRINSIC: exts.add(p, pi.id); break;
        case SPEED_BIAS: sbs.add(p, pi.id); break;
        case LANDMARK: lms.add(p, pi.id); break;
        case TD:
          if (m.has_td && m.td_ptr != p) { err = "more than one td block in one problem"; return false; }
          m.has_td = true; m.td_ptr = p; m.td = *p; break;
        default: err = "parameter block type not handled by the BA path (POSE_4D / ROTMAT / REL_COOR belong to d2pgo)"; return false;
      }
    }
  }
  // ---- 2. the estimator's properties callback on the bookkeeping problem, constness read back
  if (func_set_properties) func_set_properties();
  else setStateProperties();
  auto constness = [&](const std::vector<double *> &ptrs, std::vector<uint8_t> &out) {
    out.clear();
    for (double *p : ptrs) out.push_back(problem->IsParameterBlockConstant(p) ? 1 : 0);
  };
  constness(m.pose_ptr, m.pose_const); constness(m.ext_ptr, m.ext_const); constness(m.sb_ptr, m.sb_const);
  if (m.has_td) m.td_const = problem->IsParameterBlockConstant(m.td_ptr) ? 1 : 0;
  // ---- 3. residuals -> flat records (residual_type switch, ids from the *ResInfo, constants from the factor)
  for (auto &ri : residuals) {
    ceres::CostFunction *cf = ri->cost_function.get();
    switch (ri->residual_type) {
      case LandmarkTwoFrameOneCamResidual: {
        auto *info = static_cast<LandmarkTwoFrameOneCamResInfo *>(ri.get());
        d2ba_proj_obs o; memset(&o, 0, sizeof o);
        o.frame_a = info->frame_ida; o.frame_b = info->frame_idb; o.landmark_id = info->landmark_id; o.cam_a = info->camera_id; o.cam_b = info->camera_id;
        if (info->enable_depth_mea) {
          auto *f = dynamic_cast<ProjectionTwoFrameOneCamDepthFactor *>(cf);
          if (!f) { err = "depth-enabled two-frame residual without a ProjectionTwoFrameOneCamDepthFactor"; return false; }
          o.type = D2BA_PROJ_2F1C_DEPTH; fill_obs(o, *f); o.depth = 1.0 / f->inv_depth_j;
        } else {
          auto *f = dynamic_cast<ProjectionTwoFrameOneCamFactor *>(cf);
          if (!f) { err = "LandmarkTwoFrameOneCamResidual without a ProjectionTwoFrameOneCamFactor"; return false; }
          o.type = D2BA_PROJ_2F1C; fill_obs(o, *f);
        }
        m.obs.push_back(o); break;
      }
      case LandmarkTwoFrameTwoCamResidual: {
        auto *info = static_cast<LandmarkTwoFrameTwoCamResInfo *>(ri.get());
        auto *f = dynamic_cast<ProjectionTwoFrameTwoCamFactor *>(cf);
        if (!f) { err = "LandmarkTwoFrameTwoCamResidual without a ProjectionTwoFrameTwoCamFactor"; return false; }
        d2ba_proj_obs o; memset(&o, 0, sizeof o);
        o.type = D2BA_PROJ_2F2C; o.frame_a = info->frame_ida; o.frame_b = info->frame_idb; o.landmark_id = info->landmark_id;
        o.cam_a = info->camera_id_a; o.cam_b = info->camera_id_b; fill_obs(o, *f);
        m.obs.push_back(o); break;
      }
      case LandmarkOneFrameTwoCamResidual: {
        auto *info = static_cast<LandmarkOneFrameTwoCamResInfo *>(ri.get());
        auto *f = dynamic_cast<ProjectionOneFrameTwoCamFactor *>(cf);
        if (!f) { err = "LandmarkOneFrameTwoCamResidual without a ProjectionOneFrameTwoCamFactor"; return false; }
        d2ba_proj_obs o; memset(&o, 0, sizeof o);
        o.type = D2BA_PROJ_1F2C; o.frame_a = info->frame_ida; o.frame_b = info->frame_ida; o.landmark_id = info->landmark_id;
        o.cam_a = info->camera_id_a; o.cam_b = info->camera_id_b; fill_obs(o, *f);
        m.obs.push_back(o); break;
      }
      case DepthResidual: {
        // OneFrameDepth sits inside a ceres::AutoDiffCostFunction (depth_factor.h:9-29): r(x) = (x - 1/depth) * s, read
        // through the public Evaluate at two points
        auto *info = static_cast<DepthResInfo *>(ri.get());
        double x0 = 0.0, x1 = 1.0, r0 = 0.0, r1 = 0.0; const double *p0[1] = {&x0}, *p1[1] = {&x1};
        cf->Evaluate(p0, &r0, nullptr); cf->Evaluate(p1, &r1, nullptr);
        const double s = r1 - r0;
        d2ba_proj_obs o; memset(&o, 0, sizeof o);
        o.type = D2BA_PROJ_DEPTH_PRIOR; o.frame_a = info->base_frame_id; o.landmark_id = info->landmark_id; o.depth = -s / r0;
        m.obs.push_back(o); break;
      }
      case IMUResidual: {
        auto *info = static_cast<ImuResInfo *>(ri.get());
        auto *f = dynamic_cast<IMUFactor *>(cf);
        if (!f) { err = "IMUResidual without an IMUFactor"; return false; }
        const IntegrationBase &pre = *f->pre_integration;
        d2ba_imu r; memset(&r, 0, sizeof r);
        r.frame_a = info->frame_ida; r.frame_b = info->frame_idb; r.sum_dt = pre.sum_dt;
        copy3(r.delta_p, pre.delta_p); copy3(r.delta_v, pre.delta_v); copy3(r.linearized_ba, pre.linearized_ba); copy3(r.linearized_bg, pre.linearized_bg);
        r.delta_q[0] = pre.delta_q.x(); r.delta_q[1] = pre.delta_q.y(); r.delta_q[2] = pre.delta_q.z(); r.delta_q[3] = pre.delta_q.w();
        for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) { r.jacobian[i * 15 + j] = pre.jacobian(i, j); r.covariance[i * 15 + j] = pre.covariance(i, j); }
        m.imu.push_back(r); break;
      }
      case PriorResidual: {
        auto *f = dynamic_cast<PriorFactor *>(cf);
        if (!f) { err = "PriorResidual without a PriorFactor"; return false; }
        if (m.prior_m > 0) { err = "more than one prior"; return false; }
        std::vector<ParamInfo> keep = f->getKeepParams();
        const int mm = f->num_residuals();
        // evaluate at the linearisation points: r = e0, Jacobian blocks = columns of J_lin (prior_factor.cpp:45-90)
        std::vector<const double *> xs; std::vector<std::vector<double>> J(keep.size()); std::vector<double *> Jp;
        int eff = 0;
        for (size_t k = 0; k < keep.size(); k++) {
          xs.push_back(keep[k].data_copied.data());
          J[k].assign((size_t)mm * keep[k].size, 0.0); Jp.push_back(J[k].data());
          eff += keep[k].eff_size;
        }
        if (eff != mm) { err = "prior: residual size differs from the kept tangent dimension"; return false; }
        m.prior_e0.assign(mm, 0.0);
        if (!cf->Evaluate(xs.data(), m.prior_e0.data(), Jp.data())) { err = "prior: Evaluate failed"; return false; }
        m.prior_m = mm; m.prior_J.assign((size_t)mm * mm, 0.0);
        int off = 0;
        for (size_t k = 0; k < keep.size(); k++) {
          d2ba_blockref ref; ref.pad = 0; ref.id = keep[k].id;
          switch (keep[k].type) {
            case POSE: ref.kind = D2BA_POSE; break;
            case EXTRINSIC: ref.kind = D2BA_EXTRINSIC; break;
            case SPEED_BIAS: ref.kind = D2BA_SPEED_BIAS; break;
            case TD: ref.kind = D2BA_TD; break;
            case LANDMARK: ref.kind = D2BA_LANDMARK; break;
            default: err = "prior: block type not handled"; return false;
          }
          m.prior_refs.push_back(ref);
          for (int q = 0; q < keep[k].size; q++) m.prior_x0.push_back(keep[k].data_copied(q));
          for (int r = 0; r < mm; r++) for (int c = 0; c < keep[k].eff_size; c++) m.prior_J[(size_t)r * mm + off + c] = J[k][(size_t)r * keep[k].size + c];
          off += keep[k].eff_size;
        }
        break;
      }
      default: err = "residual type not on the BA path"; return false;
    }
  }
  return true;
}

SolverReport D2GpuSolver::solve(std::function<void()> func_set_properties) {
  SolverReport rep;
  auto failed = [&](const std::string &what) { rep.succ = false; rep.message = what; err_ = what; return rep; };
  if (!marshal(func_set_properties, last_, err_)) return failed("D2GpuSolver::marshal: " + err_);
  if (!ensureHandle()) return failed(err_);
  D2GpuMarshalled &m = last_;
  auto ck = [&](int rc) { if (rc) err_ = d2ba_last_error(h_); return rc == 0; };
  const int64_t zero = 0;
  if (!ck(d2ba_reset(h_))) return failed("d2ba_reset: " + err_);
  if (!ck(d2ba_set_blocks(h_, 0, D2BA_POSE, (int)m.pose_ids.size(), m.pose_ids.data(), m.poses.data(), m.pose_const.data())) ||
      !ck(d2ba_set_blocks(h_, 0, D2BA_EXTRINSIC, (int)m.ext_ids.size(), m.ext_ids.data(), m.exts.data(), m.ext_const.data())) ||
      !ck(d2ba_set_blocks(h_, 0, D2BA_SPEED_BIAS, (int)m.sb_ids.size(), m.sb_ids.data(), m.sbs.data(), m.sb_const.data())) ||
      (m.has_td && !ck(d2ba_set_blocks(h_, 0, D2BA_TD, 1, &zero, &m.td, &m.td_const))) ||
      !ck(d2ba_set_blocks(h_, 0, D2BA_LANDMARK, (int)m.lm_ids.size(), m.lm_ids.data(), m.lms.data(), nullptr)))
    return failed("d2ba_set_blocks: " + err_);
  if (!m.obs.empty() && !ck(d2ba_add_proj(h_, 0, (int)m.obs.size(), m.obs.data()))) return failed("d2ba_add_proj: " + err_);
  if (!m.imu.empty() && !ck(d2ba_add_imu(h_, 0, (int)m.imu.size(), m.imu.data()))) return failed("d2ba_add_imu: " + err_);
  if (m.prior_m > 0 && !ck(d2ba_set_prior(h_, 0, m.prior_m, m.prior_J.data(), m.prior_e0.data(), (int)m.prior_refs.size(), m.prior_refs.data(), m.prior_x0.data())))
    return failed("d2ba_set_prior: " + err_);
  if (cfg_.consensus_max_steps > 0 && !cons_refs_.empty() &&
      !ck(d2ba_set_consensus(h_, 0, (int)cons_refs_.size(), cons_refs_.data(), cons_slots_.data(), cons_n_slots_)))
    return failed("d2ba_set_consensus: " + err_);
  if (!ck(d2ba_finalize(h_))) return failed("d2ba_finalize: " + err_);
  d2ba_report r;
  if (!ck(d2ba_solve(h_, &r))) return failed("d2ba_solve: " + err_);
  // ---- in-place write-back through the raw block pointers (what ceres does; syncFromState then reads them)
  std::vector<double> out;
  auto scatter = [&](int kind, const std::vector<int64_t> &ids, const std::vector<double *> &ptrs, int size) {
    if (ids.empty()) return true;
    out.assign(ids.size() * size, 0.0);
    if (!ck(d2ba_get_blocks(h_, 0, kind, (int)ids.size(), ids.data(), out.data()))) return false;
    for (size_t i = 0; i < ids.size(); i++) memcpy(ptrs[i], &out[i * size], sizeof(double) * size);
    return true;
  };
  if (!scatter(D2BA_POSE, m.pose_ids, m.pose_ptr, POSE_SIZE) || !scatter(D2BA_EXTRINSIC, m.ext_ids, m.ext_ptr, POSE_SIZE) ||
      !scatter(D2BA_SPEED_BIAS, m.sb_ids, m.sb_ptr, FRAME_SPDBIAS_SIZE) || !scatter(D2BA_LANDMARK, m.lm_ids, m.lm_ptr, INV_DEP_SIZE))
    return failed("d2ba_get_blocks: " + err_);
  if (m.has_td && !m.td_const) { double td = 0; if (!ck(d2ba_get_blocks(h_, 0, D2BA_TD, 1, &zero, &td))) return failed("d2ba_get_blocks(td): " + err_); *m.td_ptr = td; }
  // SolverReport as CeresSolver::solve fills it (SolverWrapper.cpp:40-46)
  rep.total_iterations = r.total_iterations; rep.total_time = r.total_time; rep.initial_cost = r.initial_cost; rep.final_cost = r.final_cost;
  rep.state_changes = r.state_changes; rep.succ = r.succ != 0;
  rep.summary.num_successful_steps = r.successful_steps; rep.summary.num_unsuccessful_steps = r.total_iterations - r.successful_steps;
  rep.summary.initial_cost = r.initial_cost; rep.summary.final_cost = r.final_cost; rep.summary.total_time_in_seconds = r.total_time;
  return rep;
}

// d2estimator.cpp:358-423, minus the manifold assignments (every POSE / EXTRINSIC block is SE(3) with the right-multiplicative
// retraction of PoseLocalParameterization in the CUDA solver, pose_local_parameterization.cpp:13-38)
void applyStateProperties(ceres::Problem &problem, const StatePropertyInputs &in) {
  bool is_first = true;
  for (double *p : in.extrinsics_of_self_in_order) {
    if (is_first && in.not_estimate_first_extrinsic && problem.HasParameterBlock(p)) { problem.SetParameterBlockConstant(p); is_first = false; }   // :377-381
    if (!problem.HasParameterBlock(p)) continue;
    if (!in.estimate_extrinsic || !in.window_full || !in.moving) problem.SetParameterBlockConstant(p);   // :395-400
  }
  for (double *p : in.other_extrinsics)
    if (problem.HasParameterBlock(p) && (!in.estimate_extrinsic || !in.window_full || !in.moving)) problem.SetParameterBlockConstant(p);
  if (in.td && problem.HasParameterBlock(in.td) && (!in.estimate_td || !in.window_full || !in.moving)) problem.SetParameterBlockConstant(in.td);   // :412-416
  if (in.first_pose_of_self && problem.HasParameterBlock(in.first_pose_of_self) && (!in.has_prior || in.always_fixed_first_pose))
    problem.SetParameterBlockConstant(in.first_pose_of_self);   // :418-422
}

}  // namespace D2VINS
