// TEST INFRASTRUCTURE.  The application globals the reference's factor sources read (declared extern in
// RVI/parameter/parameters.h, defined in parameters.cpp, which needs OpenCV / yaml and is not compiled).
#include "parameter/parameters.h"
Eigen::Vector3d Pbg;
Eigen::Matrix3d Rwgw;
Eigen::Vector3d G;
double ACC_N, ACC_W, GYR_N, GYR_W;
bool USE_GLOBAL_OPTIMIZATION = false;   // keeps IMUGNSSBase away from the ceres::Problem it is constructed with
double MAX_TRUST_REGION_RADIUS = 1e4;
