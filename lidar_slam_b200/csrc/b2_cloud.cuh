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
