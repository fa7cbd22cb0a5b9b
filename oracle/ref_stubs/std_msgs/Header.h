// TEST INFRASTRUCTURE stub: RVI/swf/swf.h includes the ROS header type for one member declaration.
#pragma once
namespace std_msgs {
struct Header {
  double stamp = 0;
};
}
