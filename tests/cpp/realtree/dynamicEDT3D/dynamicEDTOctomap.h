// TEST FIXTURE: DynamicEDTOctomap with the accessor the reference calls (dynamicEDT3D's dynamicEDTOctomap.h:
// void getDistanceAndClosestObstacle(const octomap::point3d& p, float& distance, octomap::point3d& closestObstacle) const,
// distance in metres, distanceValue_Error = -1 outside the map).  Backed by plain arrays the test fills.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include <octomap/OcTree.h>
class DynamicEDTOctomap {
public:
    static float distanceValue_Error;
    DynamicEDTOctomap(const std::vector<float>& dist, const std::vector<int32_t>& obst, const int32_t dims[3], const int32_t min_key[3], double res)
        : dist_(dist), obst_(obst), res_(res) { for (int k = 0; k < 3; k++) { dims_[k] = dims[k]; mk_[k] = min_key[k]; } }
    void getDistanceAndClosestObstacle(const octomap::point3d& p, float& distance, octomap::point3d& closestObstacle) const {
        int c[3];
        for (unsigned k = 0; k < 3; k++) c[k] = (int)std::floor((double)p(k) / res_) - mk_[k];
        if (c[0] < 0 || c[0] >= dims_[0] || c[1] < 0 || c[1] >= dims_[1] || c[2] < 0 || c[2] >= dims_[2]) { distance = distanceValue_Error; return; }
        const size_t i = ((size_t)c[0] * dims_[1] + c[1]) * dims_[2] + c[2];
        distance = dist_[i];
        if (obst_[3 * i] >= 0)
            closestObstacle = octomap::point3d((float)(((double)(obst_[3 * i] + mk_[0]) + 0.5) * res_), (float)(((double)(obst_[3 * i + 1] + mk_[1]) + 0.5) * res_),
                                               (float)(((double)(obst_[3 * i + 2] + mk_[2]) + 0.5) * res_));
    }
    mutable long calls = 0;
private:
    std::vector<float> dist_;
    std::vector<int32_t> obst_;
    int32_t dims_[3], mk_[3];
    double res_;
};
inline float DynamicEDTOctomap::distanceValue_Error = -1.0f;
