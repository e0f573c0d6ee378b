// Stand-in for d2frontend/d2landmark_manager.h -- oracle/_ref build only: the base class / record types that
// d2vins/src/estimator/landmark_manager.hpp names in declarations (prior_factor.cpp reaches it through marginalization.hpp).
#pragma once
#include <map>
#include <d2common/d2vinsframe.h>
namespace D2Common {
struct LandmarkPerId { LandmarkIdType landmark_id = -1; };
struct VisualImageDescArray {};
}  // namespace D2Common
namespace D2FrontEnd {
using namespace D2Common;
class LandmarkManager {
 public:
  virtual ~LandmarkManager() {}
  virtual void removeLandmark(const LandmarkIdType &) {}
  const std::map<LandmarkIdType, LandmarkPerId> &getLandmarkDB() const { return landmark_db; }
 protected:
  std::map<LandmarkIdType, LandmarkPerId> landmark_db;
};
}  // namespace D2FrontEnd
using namespace D2Common;   // the real d2frontend headers open the namespace globally
