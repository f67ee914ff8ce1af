// b2_cloud.cuh -- device-resident point cloud handle shared by the b2cloud_* / b2vf_* / b2ndt_* entry points.
#pragma once

#include "b2_common.cuh"

// One cloud in HBM: packed float4 {x, y, z, intensity}.  Handles are bound to one device and are not
// thread-safe (same contract as b2ndt / b2vf).
struct b2cloud {
    int device = 0;
    size_t n = 0;                 // points held
    b2::DevBuf pts;               // float4[capacity]
    cudaStream_t st = nullptr;    // stream of the cloud's own operations (upload, append, crop)
    b2::DevBuf scratch;           // compaction bookkeeping (tile counts)
    b2::DevBuf alt;               // second point buffer: in-place compaction writes here, then the two are swapped
    b2::PinBuf h_stage, h_small;
    float first_xy[2] = {0.f, 0.f};   // x, y of point 0 as uploaded from the host (the de-skew's start azimuth) ...
    bool first_known = false;         // ... valid only until the cloud is written by anything but b2cloud_upload
    int reserve(size_t npts) {
        if (npts * 16 + 16 <= pts.cap) return 0;
        // grow keeping the contents
        b2::DevBuf nb;
        int rc = nb.reserve(npts * 16 + 16);
        if (rc) return rc;
        if (n) {
            cudaError_t e = cudaMemcpyAsync(nb.p, pts.p, n * 16, cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { b2::set_error("b2cloud: grow copy failed: %s", cudaGetErrorString(e)); nb.release(); return B2_ERR_CUDA; }
        }
        pts.release();
        pts = nb;
        return 0;
    }
    float4 *d() const { return pts.as<float4>(); }
};

#include <cmath>
namespace b2 {
// DistortionAdjust::AdjustCloud (src/models/scan_adjust/distortion_adjust.cpp:16-69), one point:
//   rotate about z so that the scan's first point has azimuth 0, azimuth -> time inside the sweep, undo the
//   motion p' = Rz(wz t) Ry(wy t) Rx(wx t) p + v t, rotate back.  Point 0 and the 5 degree sector around azimuth 0
//   are dropped; intensity is not carried over (the reference builds fresh points).
struct DeskewArg {
    float rin[9];        // rotation by -start_orientation (row-major)
    float rout[9];       // rotation by +start_orientation
    float vel[3], rate[3];
    float scan_period;
    __device__ __forceinline__ bool operator()(const float4 p, uint32_t i, float4 &out) const {
        out = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i == 0) return false;
        const float x = rin[0] * p.x + rin[1] * p.y + rin[2] * p.z;
        const float y = rin[3] * p.x + rin[4] * p.y + rin[5] * p.z;
        const float z = rin[6] * p.x + rin[7] * p.y + rin[8] * p.z;
        float o = atan2f(y, x);
        if (o < 0.0f) o = (float)((double)o + 2.0 * 3.14159265358979323846);
        const float delete_space = (float)(5.0 * 3.14159265358979323846 / 180.0);
        if (o < delete_space || (2.0 * 3.14159265358979323846 - (double)o) < (double)delete_space) return false;
        const float t = (float)((double)fabsf(o) / (double)(float)(2.0 * 3.14159265358979323846) * (double)scan_period - (double)scan_period / 2.0);
        float sx, cx, sy, cy, sz, cz;
        sincosf(rate[0] * t, &sx, &cx);
        sincosf(rate[1] * t, &sy, &cy);
        sincosf(rate[2] * t, &sz, &cz);
        const float x1 = x, y1 = cx * y - sx * z, z1 = sx * y + cx * z;             // Rx
        const float x2 = cy * x1 + sy * z1, y2 = y1, z2 = -sy * x1 + cy * z1;       // Ry
        const float x3 = cz * x2 - sz * y2, y3 = sz * x2 + cz * y2, z3 = z2;        // Rz
        const float ax = x3 + vel[0] * t, ay = y3 + vel[1] * t, az = z3 + vel[2] * t;
        out.x = rout[0] * ax + rout[1] * ay + rout[2] * az;
        out.y = rout[3] * ax + rout[4] * ay + rout[5] * az;
        out.z = rout[6] * ax + rout[7] * ay + rout[8] * az;
        return true;
    }
};


// DistortionAdjust::SetMotionInfo + the per-scan set-up of AdjustCloud (distortion_adjust.cpp:10-36): start azimuth
// from the scan's first point, the rotation by it and its inverse, velocities rotated into that frame.
inline void make_deskew_arg(float x0, float y0, float scan_period, const double linear_velocity[3], const double angular_velocity[3],
                            DeskewArg &D) {
    const float start = atan2f(y0, x0);
    const float c = (float)std::cos((double)start), s = (float)std::sin((double)start);
    const float rot[9] = {c, -s, 0.f, s, c, 0.f, 0.f, 0.f, 1.f};          // AngleAxisf(start, UnitZ).matrix()
    const float inv[9] = {c, s, 0.f, -s, c, 0.f, 0.f, 0.f, 1.f};          // its inverse
    for (int k = 0; k < 9; ++k) { D.rin[k] = inv[k]; D.rout[k] = rot[k]; }
    const float v[3] = {(float)linear_velocity[0], (float)linear_velocity[1], (float)linear_velocity[2]};
    const float w[3] = {(float)angular_velocity[0], (float)angular_velocity[1], (float)angular_velocity[2]};
    for (int r = 0; r < 3; ++r) {                                         // velocity_ = rotate_matrix * velocity_ (:31-32)
        D.vel[r] = rot[3 * r] * v[0] + rot[3 * r + 1] * v[1] + rot[3 * r + 2] * v[2];
        D.rate[r] = rot[3 * r] * w[0] + rot[3 * r + 1] * w[1] + rot[3 * r + 2] * w[2];
    }
    D.scan_period = scan_period;
}

// Fused ingest: scan de-skew (optional) + pcl::removeNaNFromPointCloud as ONE element-wise
// operation in front of the voxel pipeline.  A dropped point (point 0 and the deleted sector of the de-skew, any
// non-finite point) is not compacted away but written as a NaN point: every consumer of a device cloud skips
// non-finite points (voxel keys, crop, target build), so filter(ingest(cloud)) equals
// filter(removeNaN(deskew(cloud))) bit for bit without a compaction pass or a host round trip for the count.
struct IngestOp {
    DeskewArg D;
    int deskew;
#ifdef __CUDACC__
    __device__ __forceinline__ bool operator()(const float4 p, uint32_t i, float4 &out) const {
        bool keep = true;
        if (deskew) keep = D(p, i, out);
        else out = p;
        keep = keep && finite3(out.x, out.y, out.z);
        if (!keep) { const float qn = __int_as_float(0x7fc00000); out = make_float4(qn, qn, qn, 0.f); }
        return keep;
    }
#endif
};
// the plain pipeline: points pass through unchanged
struct NoIngest {
#ifdef __CUDACC__
    __device__ __forceinline__ bool operator()(const float4 p, uint32_t, float4 &out) const { out = p; return finite3(p.x, p.y, p.z); }
#endif
};
}  // namespace b2
