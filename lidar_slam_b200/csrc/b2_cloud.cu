// b2_cloud.cu -- device-resident clouds for the callers either side of the registration hot path
// (SURVEY 8(f) rows 1-2): local-map assembly and map cropping without leaving HBM.
//
// Reference semantics (paths relative to /root/reference/lidar_localization/):
//   local map   = sum over key frames of pcl::transformPointCloud(frame, pose)   src/mapping/front_end/front_end.cpp:375-410
//   BoxFilter   = pcl::CropBox with min/max = origin + size, identity box pose    src/models/cloud_filter/box_filter.cpp:27-37
//                 (callers: src/matching/matching.cpp:166-183, 255-262)
// Both are HBM-streaming kernels: 16 B read + 16 B written per point kept, coalesced float4 accesses,
// order-preserving (pcl::CropBox keeps the input order; operator+= appends).
#include <cmath>

#include "b2_cloud.cuh"

namespace b2 {
int check_device(int device);

// pcl::transformPointCloud on PointXYZI: xyz through the float 4x4 (left-to-right, no contraction),
// intensity copied; non-finite points are copied unchanged (is_dense == false path).
__global__ void __launch_bounds__(256) transform_append_kernel(const float4 *__restrict__ src, uint32_t n, const float *__restrict__ Tg,
                                                               float4 *__restrict__ dst) {
    __shared__ float T[16];
    if (threadIdx.x < 16) T[threadIdx.x] = Tg[threadIdx.x];
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = __ldg(&src[i]);
        float4 o = p;
        if (finite3(p.x, p.y, p.z)) transform_f32(T, p.x, p.y, p.z, o.x, o.y, o.z);
        dst[i] = o;
    }
}

// pcl::CropBox (negative = false, identity transform): keep finite points with min <= p <= max on every axis.
constexpr int CROP_THREADS = 256;
constexpr int CROP_PER_THREAD = 8;
constexpr uint32_t CROP_TILE = CROP_THREADS * CROP_PER_THREAD;
// The compaction passes take a functor Op: bool op(point, index, out_point) -- keep the point? and what to write.
struct BoxArg {
    float mn[3], mx[3];
    __device__ __forceinline__ bool operator()(const float4 p, uint32_t, float4 &out) const {
        out = p;
        if (!finite3(p.x, p.y, p.z)) return false;
        return !(p.x < mn[0] || p.y < mn[1] || p.z < mn[2] || p.x > mx[0] || p.y > mx[1] || p.z > mx[2]);
    }
};

// pass 1: kept points per tile
template <class Op>
__global__ void __launch_bounds__(CROP_THREADS) crop_count_kernel(const float4 *__restrict__ src, uint32_t n, Op B,
                                                                  uint32_t *__restrict__ tile_cnt) {
    __shared__ uint32_t wsum[CROP_THREADS / 32];
    uint32_t c = 0;
#pragma unroll
    for (int r = 0; r < CROP_PER_THREAD; ++r) {
        const uint32_t i = blockIdx.x * CROP_TILE + r * CROP_THREADS + threadIdx.x;
        float4 o;
        if (i < n) c += B(__ldg(&src[i]), i, o) ? 1u : 0u;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < CROP_THREADS / 32; ++w) t += wsum[w];
        tile_cnt[blockIdx.x] = t;
    }
}

// pass 2 (single CTA): exclusive scan of the tile counts in place; total -> tile_cnt[ntiles]
__global__ void __launch_bounds__(1024) crop_scan_kernel(uint32_t *__restrict__ tile_cnt, uint32_t ntiles) {
    chain_sync();
    __shared__ uint32_t wt[33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t base = 0;
    for (uint32_t t0 = 0; t0 < ntiles; t0 += 1024) {
        const uint32_t t = t0 + threadIdx.x;
        const uint32_t v = (t < ntiles) ? tile_cnt[t] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
        if (lane == 31) wt[w] = incl;
        __syncthreads();
        if (w == 0) {
            const uint32_t x = wt[lane];
            uint32_t xi = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, xi, d); if (lane >= d) xi += u; }
            wt[lane] = xi - x;
            if (lane == 31) wt[32] = xi;
        }
        __syncthreads();
        if (t < ntiles) tile_cnt[t] = base + wt[w] + incl - v;
        base += wt[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_cnt[ntiles] = base;
}

// pass 3: stable scatter (input order kept): rank inside the tile = rounds before + warps before + lanes before
template <class Op>
__global__ void __launch_bounds__(CROP_THREADS) crop_scatter_kernel(const float4 *__restrict__ src, uint32_t n, Op B,
                                                                    const uint32_t *__restrict__ tile_off, float4 *__restrict__ dst) {
    chain_sync();
    __shared__ uint32_t wcnt[CROP_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t base = tile_off[blockIdx.x];
#pragma unroll 1
    for (int r = 0; r < CROP_PER_THREAD; ++r) {
        const uint32_t i = blockIdx.x * CROP_TILE + r * CROP_THREADS + threadIdx.x;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        bool keep = false;
        if (i < n) { const float4 q = __ldg(&src[i]); keep = B(q, i, p); }
        const uint32_t b = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wcnt[w] = __popc(b);
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int k = 0; k < CROP_THREADS / 32; ++k) { const uint32_t c = wcnt[k]; if (k < w) before += c; total += c; }
        if (keep) dst[base + before + __popc(b & ((1u << lane) - 1u))] = p;
        base += total;
        __syncthreads();
    }
}

}  // namespace b2

using namespace b2;

// host-side packers live in b2_voxel.cu
namespace b2 {
void pack_cloud_f4(const void *src, size_t n, size_t stride, size_t ioff, float *dst_f4);
}

// One pinned staging buffer per process for cloud uploads / downloads (cudaMallocHost costs milliseconds;
// clouds are created per frame).  Handles are not thread-safe, the pool is.
#include <mutex>
static b2::PinBuf g_stage;
static std::mutex g_stage_mu;

static int cloud_check(const char *fn, const b2cloud *c) {
    if (!c) { set_error("%s: NULL cloud handle", fn); return B2_ERR_INVALID; }
    return 0;
}

extern "C" int b2cloud_create(int device, b2cloud **out) {
    if (!out) { set_error("b2cloud_create: out is NULL"); return B2_ERR_INVALID; }
    *out = nullptr;
    int rc = check_device(device);
    if (rc) return rc;
    B2_CUDA(cudaSetDevice(device));
    b2cloud *c = new b2cloud();
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e)); delete c; return B2_ERR_CUDA; }
    *out = c;
    return 0;
}

extern "C" void b2cloud_destroy(b2cloud *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->st);
    c->pts.release(); c->alt.release(); c->scratch.release(); c->h_stage.release(); c->h_small.release();
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

extern "C" int b2cloud_size(b2cloud *c, size_t *n) {
    int rc = cloud_check("b2cloud_size", c);
    if (rc) return rc;
    if (!n) { set_error("b2cloud_size: n is NULL"); return B2_ERR_INVALID; }
    *n = c->n;
    return 0;
}

extern "C" int b2cloud_clear(b2cloud *c) {
    int rc = cloud_check("b2cloud_clear", c);
    if (rc) return rc;
    c->n = 0;
    return 0;
}

extern "C" int b2cloud_device_ptr(b2cloud *c, void **d_f4) {
    int rc = cloud_check("b2cloud_device_ptr", c);
    if (rc) return rc;
    if (!d_f4) { set_error("b2cloud_device_ptr: NULL argument"); return B2_ERR_INVALID; }
    *d_f4 = c->pts.p;
    return 0;
}

extern "C" int b2cloud_upload(b2cloud *c, const void *pts, size_t n, size_t stride, size_t ioff) {
    int rc = cloud_check("b2cloud_upload", c);
    if (rc) return rc;
    if (n && !pts) { set_error("b2cloud_upload: NULL cloud"); return B2_ERR_INVALID; }
    if (stride < 16 || (stride & 3) || ioff + 4 > stride || (ioff & 3)) { set_error("b2cloud_upload: bad stride %zu / intensity offset %zu", stride, ioff); return B2_ERR_INVALID; }
    if (n >= 0xFFFFFFF0ull) { set_error("b2cloud_upload: cloud too large"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(c->device));
    c->n = 0;
    c->first_known = false;
    if ((rc = c->reserve(n))) return rc;
    if (n) {
        std::lock_guard<std::mutex> lk(g_stage_mu);
        if ((rc = g_stage.reserve(n * 16))) return rc;
        pack_cloud_f4(pts, n, stride, ioff, g_stage.as<float>());
        B2_CUDA(cudaMemcpyAsync(c->pts.p, g_stage.p, n * 16, cudaMemcpyHostToDevice, c->st));
        B2_CUDA(cudaStreamSynchronize(c->st));
    }
    c->n = n;
    if (n) {     // the de-skew needs the azimuth of the scan's first point: known here, so no device read later
        const char *q = (const char *)pts;
        c->first_xy[0] = ((const float *)q)[0]; c->first_xy[1] = ((const float *)q)[1];
        c->first_known = true;
    }
    return 0;
}

extern "C" int b2cloud_download(b2cloud *c, void *out, size_t capacity, size_t stride, size_t ioff, size_t *n) {
    int rc = cloud_check("b2cloud_download", c);
    if (rc) return rc;
    if (!n) { set_error("b2cloud_download: n is NULL"); return B2_ERR_INVALID; }
    *n = c->n;
    if (c->n == 0) return 0;
    if (!out) { set_error("b2cloud_download: NULL output"); return B2_ERR_INVALID; }
    if (stride < 16 || (stride & 3) || ioff + 4 > stride || (ioff & 3)) { set_error("b2cloud_download: bad stride / intensity offset"); return B2_ERR_INVALID; }
    if (capacity < c->n) { set_error("b2cloud_download: capacity %zu < %zu points", capacity, c->n); return B2_ERR_CAPACITY; }
    B2_CUDA(cudaSetDevice(c->device));
    std::lock_guard<std::mutex> lk(g_stage_mu);
    if ((rc = g_stage.reserve(c->n * 16))) return rc;
    B2_CUDA(cudaMemcpyAsync(g_stage.p, c->pts.p, c->n * 16, cudaMemcpyDeviceToHost, c->st));
    B2_CUDA(cudaStreamSynchronize(c->st));
    const float *res = g_stage.as<float>();
    if (stride == 16 && ioff == 12) { memcpy(out, res, c->n * 16); return 0; }
    char *o = (char *)out;
    for (size_t j = 0; j < c->n; ++j) {
        float *q = (float *)(o + j * stride);
        if (stride >= 32) memset(q, 0, stride);
        q[0] = res[4 * j]; q[1] = res[4 * j + 1]; q[2] = res[4 * j + 2];
        if (stride >= 32 || ioff != 12) q[3] = 1.0f;     // PointXYZI padding data[3]
        *(float *)(o + j * stride + ioff) = res[4 * j + 3];
    }
    return 0;
}

extern "C" int b2cloud_append_transformed(b2cloud *dst, b2cloud *src, const float T[16]) {
    int rc = cloud_check("b2cloud_append_transformed", dst);
    if (rc) return rc;
    if ((rc = cloud_check("b2cloud_append_transformed", src))) return rc;
    if (!T) { set_error("b2cloud_append_transformed: NULL transform"); return B2_ERR_INVALID; }
    if (dst == src) { set_error("b2cloud_append_transformed: dst == src"); return B2_ERR_INVALID; }
    if (dst->device != src->device) { set_error("b2cloud_append_transformed: clouds live on different devices"); return B2_ERR_INVALID; }
    if (src->n == 0) return 0;
    if (dst->n + src->n >= 0xFFFFFFF0ull) { set_error("b2cloud_append_transformed: cloud too large"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(dst->device));
    if ((rc = dst->reserve(dst->n + src->n))) return rc;
    if ((rc = dst->h_small.reserve(256))) return rc;
    if ((rc = dst->scratch.reserve(256))) return rc;
    memcpy(dst->h_small.p, T, 64);
    B2_CUDA(cudaMemcpyAsync(dst->scratch.p, dst->h_small.p, 64, cudaMemcpyHostToDevice, dst->st));
    unsigned blocks = (unsigned)((src->n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    transform_append_kernel<<<blocks, 256, 0, dst->st>>>(src->d(), (uint32_t)src->n, dst->scratch.as<float>(), dst->d() + dst->n);
    B2_LAUNCH_CHECK();
    B2_CUDA(cudaStreamSynchronize(dst->st));
    dst->n += src->n;
    dst->first_known = false;
    return 0;
}

// Local-map assembly of the front end (front_end.cpp:398-407: for every key frame  pcl::transformPointCloud + operator+=)
// in ONE launch: blockIdx.y = key frame, each with its own pose; dst = the concatenation in key-frame order.
struct AsmDesc { const float4 *src; uint32_t n, off; float T[16]; };
__global__ void __launch_bounds__(256) assemble_kernel(const AsmDesc *__restrict__ descs, float4 *__restrict__ dst) {
    __shared__ AsmDesc D;
    if (threadIdx.x == 0) D = descs[blockIdx.y];
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < D.n; i += gridDim.x * blockDim.x) {
        const float4 p = __ldg(&D.src[i]);
        float4 o = p;
        if (finite3(p.x, p.y, p.z)) transform_f32(D.T, p.x, p.y, p.z, o.x, o.y, o.z);
        dst[D.off + i] = o;
    }
}

extern "C" int b2cloud_assemble(b2cloud *dst, b2cloud *const *srcs, const float *poses, size_t K) {
    int rc = cloud_check("b2cloud_assemble", dst);
    if (rc) return rc;
    if (K && (!srcs || !poses)) { set_error("b2cloud_assemble: NULL argument"); return B2_ERR_INVALID; }
    if (K > 4096) { set_error("b2cloud_assemble: at most 4096 clouds per call"); return B2_ERR_INVALID; }
    size_t total = 0, nmax = 0;
    for (size_t k = 0; k < K; ++k) {
        if ((rc = cloud_check("b2cloud_assemble", srcs[k]))) return rc;
        if (srcs[k] == dst) { set_error("b2cloud_assemble: a source is the destination"); return B2_ERR_INVALID; }
        if (srcs[k]->device != dst->device) { set_error("b2cloud_assemble: clouds live on different devices"); return B2_ERR_INVALID; }
        total += srcs[k]->n;
        if (srcs[k]->n > nmax) nmax = srcs[k]->n;
    }
    if (total >= 0xFFFFFFF0ull) { set_error("b2cloud_assemble: cloud too large"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(dst->device));
    dst->n = 0; dst->first_known = false;
    if (total == 0) return 0;
    if ((rc = dst->reserve(total))) return rc;
    if ((rc = dst->h_small.reserve(K * sizeof(AsmDesc) + 256))) return rc;
    if ((rc = dst->scratch.reserve(K * sizeof(AsmDesc) + 256))) return rc;
    AsmDesc *hd = dst->h_small.as<AsmDesc>();
    uint32_t off = 0;
    for (size_t k = 0; k < K; ++k) {
        hd[k].src = srcs[k]->d(); hd[k].n = (uint32_t)srcs[k]->n; hd[k].off = off;
        memcpy(hd[k].T, poses + 16 * k, 64);
        off += (uint32_t)srcs[k]->n;
    }
    B2_CUDA(cudaMemcpyAsync(dst->scratch.p, hd, K * sizeof(AsmDesc), cudaMemcpyHostToDevice, dst->st));
    unsigned bx = (unsigned)((nmax + 255) / 256);
    const unsigned cap = (unsigned)((148 * 16 + K - 1) / K);          // ~16 CTAs per SM over all key frames
    if (bx > cap) bx = cap < 1 ? 1 : cap;
    assemble_kernel<<<dim3(bx, (unsigned)K), 256, 0, dst->st>>>(dst->scratch.as<AsmDesc>(), dst->d());
    B2_LAUNCH_CHECK();
    B2_CUDA(cudaStreamSynchronize(dst->st));
    dst->n = total;
    return 0;
}

// order-preserving compaction of src into dst under Op (count per tile, scan, stable scatter)
template <class Op>
static int compact_cloud(const char *fn, b2cloud *src, b2cloud *dst, const Op &B) {
    int rc = cloud_check(fn, src);
    if (rc) return rc;
    if ((rc = cloud_check(fn, dst))) return rc;
    if (dst->device != src->device) { set_error("%s: clouds live on different devices", fn); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(dst->device));
    // dst == src (the reference calls its filters with in == out, e.g. matching.cpp:158): the survivors go to the
    // cloud's second buffer and the two buffers are swapped, so the cloud's device pointer changes
    const bool in_place = (dst == src);
    const size_t n = src->n;
    if (!in_place) dst->n = 0;
    if (n == 0) return 0;
    const uint32_t ntiles = (uint32_t)((n + CROP_TILE - 1) / CROP_TILE);
    if (in_place) { if ((rc = dst->alt.reserve(n * 16 + 16))) return rc; }
    else if ((rc = dst->reserve(n))) return rc;
    if ((rc = dst->scratch.reserve((size_t)(ntiles + 1) * 4 + 256))) return rc;
    if ((rc = dst->h_small.reserve(256))) return rc;
    uint32_t *tiles = dst->scratch.as<uint32_t>() + 64;     // first 256 bytes hold the transform of append
    float4 *out = in_place ? dst->alt.as<float4>() : dst->d();
    crop_count_kernel<<<ntiles, CROP_THREADS, 0, dst->st>>>(src->d(), (uint32_t)n, B, tiles);
    B2_LAUNCH_CHECK();
    launch_chain(crop_scan_kernel, 1, 1024, 0, dst->st, tiles, ntiles);
    B2_LAUNCH_CHECK();
    launch_chain(crop_scatter_kernel<Op>, ntiles, CROP_THREADS, 0, dst->st, (const float4 *)src->d(), (uint32_t)n, B, (const uint32_t *)tiles, out);
    B2_LAUNCH_CHECK();
    B2_CUDA(cudaMemcpyAsync(dst->h_small.p, tiles + ntiles, 4, cudaMemcpyDeviceToHost, dst->st));
    B2_CUDA(cudaStreamSynchronize(dst->st));
    if (in_place) { const DevBuf t = dst->pts; dst->pts = dst->alt; dst->alt = t; }
    dst->n = *dst->h_small.as<uint32_t>();
    dst->first_known = false;
    return 0;
}

extern "C" int b2cloud_box_filter(b2cloud *src, const float edge[6], b2cloud *dst) {
    if (!edge) { set_error("b2cloud_box_filter: NULL edge"); return B2_ERR_INVALID; }
    BoxArg B;
    for (int a = 0; a < 3; ++a) { B.mn[a] = edge[2 * a]; B.mx[a] = edge[2 * a + 1]; }
    return compact_cloud("b2cloud_box_filter", src, dst, B);
}

// DistortionAdjust::SetMotionInfo + AdjustCloud (distortion_adjust.cpp:10-69): scan_period in seconds, velocities of
// the sensor in its own frame (VelocityData linear / angular, double in the reference, used as float).
extern "C" int b2cloud_distortion_adjust(b2cloud *src, float scan_period, const double linear_velocity[3],
                                         const double angular_velocity[3], b2cloud *dst) {
    if (!src || !dst || !linear_velocity || !angular_velocity) { set_error("b2cloud_distortion_adjust: NULL argument"); return B2_ERR_INVALID; }
    if (src->n == 0) { dst->n = 0; return 0; }     // in == out is allowed (data_pretreat_flow.cpp calls AdjustCloud that way)
    B2_CUDA(cudaSetDevice(src->device));
    int rc;
    if ((rc = src->h_small.reserve(256))) return rc;
    // start_orientation = atan2(points[0].y, points[0].x)
    B2_CUDA(cudaMemcpyAsync(src->h_small.p, src->pts.p, 16, cudaMemcpyDeviceToHost, src->st));
    B2_CUDA(cudaStreamSynchronize(src->st));
    const float *p0 = src->h_small.as<float>();
    DeskewArg D;
    make_deskew_arg(p0[0], p0[1], scan_period, linear_velocity, angular_velocity, D);
    return compact_cloud("b2cloud_distortion_adjust", src, dst, D);
}

// pcl::removeNaNFromPointCloud (front_end.cpp:92, matching.cpp:188): dst = the points of src whose x, y and z are
// finite, input order kept.  Same compaction kernels as the crop with an unbounded box.
extern "C" int b2cloud_remove_nan(b2cloud *src, b2cloud *dst) {
    const float inf = __builtin_huge_valf();
    const float edge[6] = {-inf, inf, -inf, inf, -inf, inf};
    return b2cloud_box_filter(src, edge, dst);
}
