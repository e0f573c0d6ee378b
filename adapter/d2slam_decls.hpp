// d2slam_decls.hpp -- the D2SLAM declarations the solver adapter compiles against, for machines without ROS / OpenCV /
// LCM / swarm_msgs (this image).  On a real D2SLAM checkout the adapter includes the reference's own headers instead
// (define D2GPU_WITH_D2SLAM_HEADERS); here the include chain of d2common/d2state.hpp (d2vinsframe.h -> d2frontend_types.h ->
// ROS messages, LCM types, cv::Mat) cannot be satisfied, so this header re-declares -- same names, same members, same
// virtual signatures -- exactly the part of the interface the solver sees:
//
//   D2Common::ParamsType / ParamInfo / ResidualType / ResidualInfo   d2common/include/d2common/solver/BaseParamResInfo.hpp:7-79
//   D2Common::SolverReport / SolverWrapper                           d2common/include/d2common/solver/SolverWrapper.hpp:14-53
//   D2VINS::*ResInfo (ids of the blocks of each residual type)       d2vins/src/estimator/ParamResidualInfo.hpp:18-189
//   D2VINS::PriorFactor public interface                             d2vins/src/factors/prior_factor.h:20-58
//   D2Common::D2State / D2VINS::D2EstimatorState state getters        d2common/include/d2common/d2state.hpp:87-110,
//                                                                    d2vins/src/estimator/d2vinsstate.hpp:55-61
//
// The FACTOR classes are not re-declared: the adapter includes the reference's own projection*Factor.h / imu_factor.h
// (compiled against oracle/_shim).  Nothing here has behaviour beyond trivial storage.
#pragma once
#include <ceres/ceres.h>
#include <d2common/d2basetypes.h>

#include <functional>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

namespace D2Common {

// ---- BaseParamResInfo.hpp:7-40
enum ParamsType { POSE = 0, POSE_4D, POSE_PERTURB_6D, ROTMAT, REL_COOR, SPEED_BIAS, EXTRINSIC, TD, LANDMARK };
inline bool IsSE3(ParamsType type) { return type == POSE || type == REL_COOR || type == EXTRINSIC; }

struct ParamInfo {
  StatePtr pointer = nullptr;
  Eigen::Matrix<state_type, -1, 1> data_copied;
  int index = -1;
  int size = 0;
  int eff_size = 0;
  bool is_remove = false;
  ParamsType type;
  FrameIdType id;
  ParamInfo() {}
  state_type *getPointer() { return CheckGetPtr(pointer); }
};

// ---- d2state.hpp (getters only)
class D2State {
 protected:
  int self_id;
  std::map<FrameIdType, StatePtr> _frame_pose_state;

 public:
  explicit D2State(int _self_id) : self_id(_self_id) {}
  virtual ~D2State() {}
  StatePtr getPoseState(FrameIdType frame_id) const {
    auto it = _frame_pose_state.find(frame_id);
    return it == _frame_pose_state.end() ? nullptr : it->second;
  }
};

// ---- BaseParamResInfo.hpp:42-86
enum ResidualType {
  NONE, IMUResidual, LandmarkTwoFrameOneCamResidual, LandmarkTwoFrameTwoCamResidual, LandmarkTwoDroneTwoCamResidual,
  LandmarkOneFrameTwoCamResidual, PriorResidual, DepthResidual, RelPoseResidual, RelRotResidual, GravityPriorResidual
};

class ResidualInfo {
 public:
  ResidualType residual_type;
  std::shared_ptr<ceres::CostFunction> cost_function = nullptr;
  std::shared_ptr<ceres::LossFunction> loss_function = nullptr;
  ResidualInfo(ResidualType type) : residual_type(type) {}
  virtual bool relavant(const std::set<FrameIdType> &frame_id) const = 0;
  virtual std::vector<ParamInfo> paramsList(D2State *state) const = 0;
  virtual std::vector<state_type *> paramsPointerList(D2State *state) const {
    std::vector<state_type *> params;
    for (auto info : paramsList(state)) params.push_back(CheckGetPtr(info.pointer));
    return params;
  }
  int residualSize() const { return cost_function->num_residuals(); }
  virtual ~ResidualInfo() {}
};
using ResidualInfoPtr = std::shared_ptr<ResidualInfo>;

inline ParamInfo createFramePose(D2State *state, FrameIdType id, bool is_perturb = false) {
  (void)is_perturb;
  ParamInfo info;
  info.type = POSE; info.pointer = state->getPoseState(id); info.size = POSE_SIZE; info.eff_size = POSE_EFF_SIZE; info.id = id;
  return info;
}

// ---- SolverWrapper.hpp:14-53
struct SolverReport {
  int total_iterations = 0;
  double total_time = 0;
  double initial_cost = 0;
  double final_cost = 0;
  double state_changes = 0;
  bool succ = true;
  std::string message = "";
  ceres::Solver::Summary summary;
};

class SolverWrapper {
 protected:
  ceres::Problem *problem = nullptr;
  ceres::Problem::Options problem_options;
  D2State *state;
  std::vector<std::shared_ptr<ResidualInfo>> residuals;
  virtual void setStateProperties() {}

 public:
  SolverWrapper(D2State *_state) : state(_state) { problem = new ceres::Problem(problem_options); }
  virtual void addResidual(const std::shared_ptr<ResidualInfo> &residual_info) { residuals.push_back(residual_info); }
  virtual SolverReport solve() = 0;
  virtual SolverReport solve(std::function<void()> func_set_properties) = 0;
  ceres::Problem &getProblem() { return *problem; }
  virtual void reset() { delete problem; problem = new ceres::Problem(problem_options); residuals.clear(); }
};
}  // namespace D2Common

namespace D2VINS {
using namespace D2Common;

// ---- prior_factor.h:20-58 (public interface; the real class keeps linearized_jac / linearized_res private)
class PriorFactor : public ceres::CostFunction {
 public:
  virtual std::vector<ParamInfo> getKeepParams() const = 0;
  virtual int getEffParamsDim() const = 0;
};
using PriorFactorPtr = std::shared_ptr<PriorFactor>;

// ---- d2vinsstate.hpp:55-61 (getters the residual descriptors use)
class D2EstimatorState : public D2State {
 protected:
  std::map<int, StatePtr> _camera_extrinsic_state;
  std::map<FrameIdType, StatePtr> spd_bias_state;
  std::map<LandmarkIdType, StatePtr> landmark_state;
  std::map<int, StatePtr> td_state;
  PriorFactorPtr prior_factor = nullptr;

 public:
  explicit D2EstimatorState(int _self_id) : D2State(_self_id) {}
  StatePtr getExtrinsicState(int i) const { auto it = _camera_extrinsic_state.find(i); return it == _camera_extrinsic_state.end() ? nullptr : it->second; }
  StatePtr getSpdBiasState(FrameIdType frame_id) const { auto it = spd_bias_state.find(frame_id); return it == spd_bias_state.end() ? nullptr : it->second; }
  StatePtr getLandmarkState(LandmarkIdType landmark_id) const { auto it = landmark_state.find(landmark_id); return it == landmark_state.end() ? nullptr : it->second; }
  StatePtr getTdState(int drone_id) { auto it = td_state.find(drone_id); return it == td_state.end() ? nullptr : it->second; }
  PriorFactorPtr getPrior() const { return prior_factor; }
};

// ---- ParamResidualInfo.cpp:29-74
inline ParamInfo createExtrinsic(D2EstimatorState *state, int camera_id) {
  ParamInfo info; info.pointer = state->getExtrinsicState(camera_id); info.size = POSE_SIZE; info.eff_size = POSE_EFF_SIZE; info.type = EXTRINSIC; info.id = camera_id; return info;
}
inline ParamInfo createLandmark(D2EstimatorState *state, int landmark_id, bool inv_dep_param = true) {
  (void)inv_dep_param;
  ParamInfo info; info.pointer = state->getLandmarkState(landmark_id); info.size = INV_DEP_SIZE; info.eff_size = INV_DEP_SIZE; info.type = LANDMARK; info.id = landmark_id; return info;
}
inline ParamInfo createSpeedBias(D2EstimatorState *state, FrameIdType id) {
  ParamInfo info; info.pointer = state->getSpdBiasState(id); info.size = FRAME_SPDBIAS_SIZE; info.eff_size = FRAME_SPDBIAS_SIZE; info.type = SPEED_BIAS; info.id = id; return info;
}
inline ParamInfo createTd(D2EstimatorState *state, int camera_id) {
  ParamInfo info; info.pointer = state->getTdState(camera_id); info.size = TD_SIZE; info.eff_size = TD_SIZE; info.type = TD; info.id = camera_id; return info;
}

// ---- ParamResidualInfo.hpp:18-189: which blocks (by id) each residual type touches, in the factor's parameter order
class LandmarkTwoFrameOneCamResInfo : public ResidualInfo {
 public:
  FrameIdType frame_ida, frame_idb; LandmarkIdType landmark_id; int camera_id; bool enable_depth_mea = false;
  LandmarkTwoFrameOneCamResInfo() : ResidualInfo(ResidualType::LandmarkTwoFrameOneCamResidual) {}
  bool relavant(const std::set<FrameIdType> &f) const override { return f.count(frame_ida) || f.count(frame_idb); }
  std::vector<ParamInfo> paramsList(D2State *state) const override {
    auto s = static_cast<D2EstimatorState *>(state);
    return {createFramePose(s, frame_ida), createFramePose(s, frame_idb), createExtrinsic(s, camera_id), createLandmark(s, landmark_id), createTd(s, camera_id)};
  }
};
class LandmarkTwoFrameTwoCamResInfo : public ResidualInfo {
 public:
  FrameIdType frame_ida, frame_idb; LandmarkIdType landmark_id; int camera_id_a, camera_id_b;
  LandmarkTwoFrameTwoCamResInfo() : ResidualInfo(ResidualType::LandmarkTwoFrameTwoCamResidual) {}
  bool relavant(const std::set<FrameIdType> &f) const override { return f.count(frame_ida) || f.count(frame_idb); }
  std::vector<ParamInfo> paramsList(D2State *state) const override {
    auto s = static_cast<D2EstimatorState *>(state);
    return {createFramePose(s, frame_ida), createFramePose(s, frame_idb), createExtrinsic(s, camera_id_a), createExtrinsic(s, camera_id_b), createLandmark(s, landmark_id), createTd(s, camera_id_a)};
  }
};
class LandmarkOneFrameTwoCamResInfo : public ResidualInfo {
 public:
  FrameIdType frame_ida; LandmarkIdType landmark_id; int camera_id_a, camera_id_b;
  LandmarkOneFrameTwoCamResInfo() : ResidualInfo(ResidualType::LandmarkOneFrameTwoCamResidual) {}
  bool relavant(const std::set<FrameIdType> &f) const override { return f.count(frame_ida) != 0; }
  std::vector<ParamInfo> paramsList(D2State *state) const override {
    auto s = static_cast<D2EstimatorState *>(state);
    return {createExtrinsic(s, camera_id_a), createExtrinsic(s, camera_id_b), createLandmark(s, landmark_id), createTd(s, camera_id_a)};
  }
};
class ImuResInfo : public ResidualInfo {
 public:
  FrameIdType frame_ida, frame_idb;
  ImuResInfo() : ResidualInfo(ResidualType::IMUResidual) {}
  bool relavant(const std::set<FrameIdType> &f) const override { return f.count(frame_ida) || f.count(frame_idb); }
  std::vector<ParamInfo> paramsList(D2State *state) const override {
    auto s = static_cast<D2EstimatorState *>(state);
    return {createFramePose(s, frame_ida), createSpeedBias(s, frame_ida), createFramePose(s, frame_idb), createSpeedBias(s, frame_idb)};
  }
};
class DepthResInfo : public ResidualInfo {
 public:
  FrameIdType base_frame_id; LandmarkIdType landmark_id;
  DepthResInfo() : ResidualInfo(ResidualType::DepthResidual) {}
  bool relavant(const std::set<FrameIdType> &f) const override { return f.count(base_frame_id) != 0; }
  std::vector<ParamInfo> paramsList(D2State *state) const override { return {createLandmark(static_cast<D2EstimatorState *>(state), landmark_id)}; }
};
class PriorResInfo : public ResidualInfo {
  PriorFactorPtr factor;
 public:
  PriorResInfo(const PriorFactorPtr &_factor) : ResidualInfo(PriorResidual), factor(_factor) { cost_function = _factor; }
  std::vector<ParamInfo> paramsList(D2State *) const override { return factor->getKeepParams(); }
  bool relavant(const std::set<FrameIdType> &) const override { return true; }
};
}  // namespace D2VINS
