// Stub for ceres/problem.h: common_function.h only needs ceres::ResidualBlockId (mea_t member).
#pragma once
namespace ceres {
namespace internal { class ResidualBlock; }
typedef internal::ResidualBlock* ResidualBlockId;
}
