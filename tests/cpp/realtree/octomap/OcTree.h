// TEST FIXTURE: octomath::Vector3 / octomap::point3d / octomap::OcTree with the member spelling of octomap 1.9
// (float storage; x() y() z(), operator()(unsigned), arithmetic, norm / distance returning double).
#pragma once
#include <cmath>
namespace octomath {
class Vector3 {
public:
    Vector3(float x = 0.f, float y = 0.f, float z = 0.f) { data[0] = x; data[1] = y; data[2] = z; }
    float& operator()(unsigned int i) { return data[i]; }
    const float& operator()(unsigned int i) const { return data[i]; }
    float& x() { return data[0]; } float& y() { return data[1]; } float& z() { return data[2]; }
    const float& x() const { return data[0]; } const float& y() const { return data[1]; } const float& z() const { return data[2]; }
    Vector3 operator-(const Vector3& o) const { return Vector3(data[0] - o.data[0], data[1] - o.data[1], data[2] - o.data[2]); }
    Vector3 operator+(const Vector3& o) const { return Vector3(data[0] + o.data[0], data[1] + o.data[1], data[2] + o.data[2]); }
    Vector3 operator*(float s) const { return Vector3(data[0] * s, data[1] * s, data[2] * s); }
    double norm() const { return std::sqrt((double)(data[0] * data[0] + data[1] * data[1] + data[2] * data[2])); }
protected:
    float data[3];
};
}  // namespace octomath
namespace octomap {
typedef octomath::Vector3 point3d;
class OcTree {
public:
    explicit OcTree(double resolution) : res_(resolution) {}
    double getResolution() const { return res_; }
private:
    double res_;
};
}  // namespace octomap
