// TEST FIXTURE: the subset of <ros/ros.h> the hot-path class signatures mention (ros::NodeHandle, ros::Time), with the
// real API's spelling, so that include/dlsc_compat.hpp can be compiled in its real-tree mode (DLSC_COMPAT_STANDALONE off).
#pragma once
namespace ros {
class NodeHandle {};
class Time {
public:
    Time() : sec_(0) {}
    explicit Time(double t) : sec_(t) {}
    static Time now() { return Time(); }
    double toSec() const { return sec_; }
private:
    double sec_;
};
}  // namespace ros
