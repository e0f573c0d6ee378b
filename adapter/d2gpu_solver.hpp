// d2gpu_solver.hpp -- D2GpuSolver: the SolverWrapper that puts libd2ba.so (include/d2ba.h) under D2VINS::D2Estimator.
//
// It takes the place of CeresSolver / D2VINSConsensusSolver in D2Estimator::init
// (d2vins/src/estimator/d2estimator.cpp:49-54):
//
//     solver = new D2VINS::D2GpuSolver(&state, D2GpuSolverConfig::fromParams(*params));
//
// and keeps the reference's contract: addResidual(ResidualInfoPtr) collects the residual descriptors the estimator builds
// in setupImuFactors / setupLandmarkFactors / setupPriorFactor (d2estimator.cpp:687-897); solve(fn) calls fn() -- the
// estimator's setStateProperties (d2estimator.cpp:358-423) -- against a ceres::Problem that holds only the parameter
// blocks, reads the block constness back with IsParameterBlockConstant, marshals blocks and residuals into the flat C-ABI
// records, runs the CUDA solve and scatters the solved blocks back through the same raw StatePtr pointers ceres would have
// written in place (BaseParamResInfo.hpp:68-74), so D2EstimatorState::syncFromState (d2vinsstate.cpp:557-592) works
// unchanged.  reset() mirrors SolverWrapper::reset (SolverWrapper.cpp:20-24).
//
// Factor constants are read through the reference's own factor classes (public members: pts_i, pts_j, velocity_i/j,
// td_i/j of the projection factors, projectionTwoFrameOneCamFactor.h:28-31; IMUFactor::pre_integration, imu_factor.h:215).
// The prior's (J, e0) are private (prior_factor.h:23-24): they are recovered through the public interface by evaluating
// the cost function at its own linearisation points (getKeepParams()[k].data_copied): r(x0) = e0, dr/dx(x0) = J.
#pragma once
#ifdef D2GPU_WITH_D2SLAM_HEADERS
#include <d2common/solver/SolverWrapper.hpp>
#include "../estimator/ParamResidualInfo.hpp"
#include "../factors/prior_factor.h"
#else
#include "d2slam_decls.hpp"
#endif
#include <cstdint>
#include <string>
#include <vector>

#include "../include/d2ba.h"

namespace D2VINS {

struct D2GpuSolverConfig {
  int device = 0;
  int max_num_iterations = 8;         // d2vins_params.cpp:144
  double max_solver_time = 0.08;      // d2vins_params.cpp:143 (seconds; <= 0: no budget)
  int consensus_max_steps = 0;        // 0: CeresSolver semantics, > 0: ConsensusSolver (d2vins_params.cpp:151-160)
  double focal_length = 460.0, depth_sqrt_inf = 20.0, gravity_norm = 9.805;
  double rho_frame_T = 100.0, rho_frame_theta = 100.0, rho_landmark = 1.0, relaxation_alpha = 0.0;
};

// The flat records one solve hands to the C ABI (kept for inspection: tests compare them with the generator's arrays)
struct D2GpuMarshalled {
  std::vector<int64_t> pose_ids, ext_ids, sb_ids, lm_ids;
  std::vector<double> poses, exts, sbs, lms;
  std::vector<uint8_t> pose_const, ext_const, sb_const;
  std::vector<double *> pose_ptr, ext_ptr, sb_ptr, lm_ptr;
  double td = 0; uint8_t td_const = 1; double *td_ptr = nullptr; bool has_td = false;
  std::vector<d2ba_proj_obs> obs;
  std::vector<d2ba_imu> imu;
  int prior_m = 0; std::vector<double> prior_J, prior_e0, prior_x0; std::vector<d2ba_blockref> prior_refs;
};

class D2GpuSolver : public D2Common::SolverWrapper {
 public:
  D2GpuSolver(D2Common::D2State *state, const D2GpuSolverConfig &cfg);
  ~D2GpuSolver();   // (SolverWrapper has no virtual destructor, SolverWrapper.hpp:39-53)
  D2Common::SolverReport solve() override { return solve(nullptr); }
  D2Common::SolverReport solve(std::function<void()> func_set_properties) override;
  void reset() override;
  // consensus mode: blocks taking part in the averaging and their global slots (ConsensusSolver::addParam, ConsensusSolver.cpp:26-37)
  void setConsensusSlots(const std::vector<d2ba_blockref> &refs, const std::vector<int32_t> &slots, int n_slots_global);
  // stage 1 of solve(): parameter blocks into the bookkeeping ceres::Problem, fn(), constness read-back, flat records.
  // No CUDA call is made (usable on a machine without a GPU).
  bool marshal(std::function<void()> func_set_properties, D2GpuMarshalled &out, std::string &err);
  const D2GpuMarshalled &lastMarshalled() const { return last_; }
  const std::string &lastError() const { return err_; }

 private:
  D2GpuSolverConfig cfg_;
  d2ba_handle *h_ = nullptr;
  D2GpuMarshalled last_;
  std::string err_;
  std::vector<d2ba_blockref> cons_refs_; std::vector<int32_t> cons_slots_; int cons_n_slots_ = 0;
  bool ensureHandle();
};

// The block-constness rules of D2Estimator::setStateProperties (d2estimator.cpp:358-423) as a free function over the
// quantities that method reads, for callers that do not run the estimator itself (the synthetic harness, tests).
struct StatePropertyInputs {
  bool estimate_extrinsic = false, estimate_td = false, always_fixed_first_pose = false, not_estimate_first_extrinsic = false;
  bool window_full = true;         // state.size(drone) >= max_sld_win_size
  bool moving = true;              // lastFrame()->odom.vel().norm() >= estimate_extrinsic_vel_thres
  bool has_prior = true;           // state.getPrior() != nullptr
  std::vector<double *> extrinsics_of_self_in_order;   // getAvailableCameraIds() order
  std::vector<double *> other_extrinsics;
  double *td = nullptr;
  double *first_pose_of_self = nullptr;
};
void applyStateProperties(ceres::Problem &problem, const StatePropertyInputs &in);

}  // namespace D2VINS
