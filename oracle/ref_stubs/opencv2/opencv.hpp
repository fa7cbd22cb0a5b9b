// TEST INFRASTRUCTURE stub: the reference's headers name a few OpenCV types in declarations (RVI/parameter/parameters.h,
// RVI/feature/*.h, RVI/swf/swf.h); none of the code compiled into oracle/_ref touches an image.  Empty stand-ins.
#pragma once
#include <iostream>
#include <list>
#include <map>
#include <queue>
#include <set>
#include <string>
#include <vector>
namespace cv {
struct Mat {};
struct Point2f {
  float x = 0, y = 0;
};
struct Point3f {
  float x = 0, y = 0, z = 0;
};
struct Scalar {};
struct Size {};
template <typename T>
struct Ptr {};
namespace cuda {
struct GpuMat {};
}
}  // namespace cv
typedef unsigned char uchar;
