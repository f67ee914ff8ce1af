// deskew_ref.cpp -- TEST INFRASTRUCTURE ONLY.  C entry point around the reference's OWN DistortionAdjust
// (lidar_localization/src/models/scan_adjust/distortion_adjust.cpp, compiled where it lies by oracle/Makefile
// against the header stand-ins of oracle/ref_stubs and the vendored Eigen 3.2.92): pins the oracle's restatement.
#include "lidar_localization/models/scan_adjust/distortion_adjust.hpp"

using namespace lidar_localization;

extern "C" int ref_distortion_adjust(const float *xyzi, int n, float scan_period, const double lin[3], const double ang[3],
                                     float *out_xyzi, int out_capacity) {
    CloudData::CLOUD_PTR in(new CloudData::CLOUD());
    for (int i = 0; i < n; ++i) {
        CloudData::POINT p;
        p.x = xyzi[4 * i]; p.y = xyzi[4 * i + 1]; p.z = xyzi[4 * i + 2]; p.intensity = xyzi[4 * i + 3];
        in->points.push_back(p);
    }
    VelocityData v;
    v.linear_velocity.x = lin[0]; v.linear_velocity.y = lin[1]; v.linear_velocity.z = lin[2];
    v.angular_velocity.x = ang[0]; v.angular_velocity.y = ang[1]; v.angular_velocity.z = ang[2];
    DistortionAdjust da;
    da.SetMotionInfo(scan_period, v);
    CloudData::CLOUD_PTR out;
    da.AdjustCloud(in, out);
    const int m = (int)out->points.size();
    for (int i = 0; i < m && i < out_capacity; ++i) {
        out_xyzi[4 * i] = out->points[i].x; out_xyzi[4 * i + 1] = out->points[i].y;
        out_xyzi[4 * i + 2] = out->points[i].z; out_xyzi[4 * i + 3] = out->points[i].intensity;
    }
    return m;
}
