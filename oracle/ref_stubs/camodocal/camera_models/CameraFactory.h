// TEST INFRASTRUCTURE stub: RVI/feature/feature_tracker.h declares members of these types; nothing compiled into
// oracle/_ref uses a camera model.
#pragma once
#include <memory>
namespace camodocal {
struct Camera {};
typedef std::shared_ptr<Camera> CameraPtr;
}  // namespace camodocal
