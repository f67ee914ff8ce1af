// b2_hmap.cu -- initial-yaw search of the matching node, device resident.
//
// Reference semantics (paths relative to /root/reference/lidar_localization/):
//   Matching::generateGauss2DMapCells  src/matching/matching.cpp:344-394  2-D grid over the local map (minus its
//       origin): per cell the running mean / variance of z, updated point by point in input order
//   Matching::resetMapRange            src/matching/matching.cpp:396-424  float min / max of the centred points
//   Matching::getInitialYawAngle       src/matching/matching.cpp:267-308  270 yaw bins: rotate the scan about z,
//       look every point up in the grid, probs[bin] += exp(-(z - mu)^2 / (2 sigma)); best = first maximal bin
// The grid build needs the reference's per-cell input order (the recurrence is not associative): points are
// sorted by cell with the stable radix sort of the voxel pipeline and one thread walks each cell.  The yaw
// search is one CTA per bin streaming the scan (HBM/L2 streaming + a gather into the grid).
#include <cfloat>
#include <cmath>
#include <vector>

#include "b2_cloud.cuh"
#include "b2_voxel.cuh"

namespace b2 {
int check_device(int device);

struct HmapGrid {
    float min_x, min_y;     // map_min_xyz_(0), (1)
    double res;             // grid_map_resolution_ (double in the reference)
    int32_t width, height;  // local_map_width_, local_map_height_
};

// resetMapRange: float min / max of (point - origin) over the finite points
__global__ void __launch_bounds__(256) hmap_bounds_kernel(const float4 *__restrict__ pts, uint32_t n, float ox, float oy, float oz,
                                                          uint32_t *__restrict__ bounds /*6 ordered floats*/) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = __ldg(&pts[i]);
        if (!finite3(p.x, p.y, p.z)) continue;
        const float c[3] = {__fsub_rn(p.x, ox), __fsub_rn(p.y, oy), __fsub_rn(p.z, oz)};
#pragma unroll
        for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], c[a]); mx[a] = fmaxf(mx[a], c[a]); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&bounds[a], f2ord(mn[a]));
            atomicMax(&bounds[3 + a], f2ord(mx[a]));
        }
    }
}

// cell of a centred coordinate: (int) round((c - min) / res), float subtraction, double division (matching.cpp:370)
__device__ __forceinline__ int hmap_cell(float c, float mn, double res) {
    return (int)round(__ddiv_rn((double)__fsub_rn(c, mn), res));
}

__global__ void __launch_bounds__(256) hmap_key_kernel(const float4 *__restrict__ pts, uint32_t n, float ox, float oy, HmapGrid G,
                                                       uint32_t *__restrict__ keys, uint32_t *__restrict__ n_valid) {
    uint32_t cnt = 0;
    const uint32_t invalid = (uint32_t)G.width * (uint32_t)G.height;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = __ldg(&pts[i]);
        uint32_t key = invalid;
        if (finite3(p.x, p.y, p.z)) {
            const int cx = hmap_cell(__fsub_rn(p.x, ox), G.min_x, G.res);
            const int cy = hmap_cell(__fsub_rn(p.y, oy), G.min_y, G.res);
            if (!(cx < 0 || cy < 0 || cx >= G.width || cy >= G.height)) { key = (uint32_t)cx * (uint32_t)G.height + (uint32_t)cy; ++cnt; }
        }
        keys[i] = key;          // the payload of the sort is the element index itself (first radix pass)
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_valid, cnt);
}

// one thread per occupied cell: the reference's recurrence over the cell's points in input order (matching.cpp:374-392)
__global__ void __launch_bounds__(128) hmap_cell_kernel(const float4 *__restrict__ pts, SortView sv, const uint32_t *__restrict__ run_start,
                                                        const uint32_t *__restrict__ n_valid,
                                                        float oz, float *__restrict__ mu_out, float *__restrict__ sigma_out,
                                                        int32_t *__restrict__ cnt_out) {
    const uint32_t runs = sv.scalars[1];
    const uint32_t *__restrict__ keys = sv.keys();
    const uint32_t *__restrict__ vals = sv.vals();
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= runs) return;
    const uint32_t s = run_start[j];
    const uint32_t e = (j + 1 < runs) ? run_start[j + 1] : *n_valid;
    float mu = 0.f, sigma = 0.f;
    int cnt = 0;
    for (uint32_t k = s; k < e; ++k) {
        const float z = __fsub_rn(__ldg(&pts[vals[k]]).z, oz);
        if (cnt == 0) {
            mu = z; sigma = 0.f; cnt = 1;
        } else {
            // mu' = (cnt * mu + z) / (cnt + 1)                                   (float)
            const float mu_n = __fdiv_rn(__fadd_rn(__fmul_rn((float)cnt, mu), z), (float)(cnt + 1));
            // sigma' = ((cnt-1) sigma + pow(z - mu, 2) + (cnt+1) pow(mu' - mu, 2) + 2 (mu' - mu)(z - mu')) / cnt
            // std::pow(float, int) is double: the sum is formed left to right in double, stored to float, then /= cnt
            const double t0 = (double)__fmul_rn((float)(cnt - 1), sigma);
            const double dz = (double)__fsub_rn(z, mu);
            const double dm = (double)__fsub_rn(mu_n, mu);
            const double t1 = __dmul_rn(dz, dz);
            const double t2 = __dmul_rn((double)(cnt + 1), __dmul_rn(dm, dm));
            const float t3 = __fmul_rn(__fmul_rn(2.0f, __fsub_rn(mu_n, mu)), __fsub_rn(z, mu_n));
            const double sum = __dadd_rn(__dadd_rn(__dadd_rn(t0, t1), t2), (double)t3);
            sigma = __fdiv_rn((float)sum, (float)cnt);
            mu = mu_n;
            ++cnt;
        }
    }
    const uint32_t cell = keys[s];
    mu_out[cell] = mu; sigma_out[cell] = sigma; cnt_out[cell] = cnt;
}

// getInitialYawAngle: one CTA per yaw bin; per-thread partial sums in double, combined in a fixed order
constexpr int YAW_THREADS = 256;
struct YawRot { float c, s, m22, tx, ty; };   // Eigen::AngleAxisf(angle, UnitZ).matrix(): [c -s 0; s c 0; 0 0 (1-c)+c]; translation (tx, ty, 0)

__global__ void __launch_bounds__(YAW_THREADS) hmap_yaw_kernel(const float4 *__restrict__ scan, uint32_t n, HmapGrid G,
                                                              const float *__restrict__ mu, const float *__restrict__ sigma,
                                                              const int32_t *__restrict__ cnt, const YawRot *__restrict__ rot,
                                                              double *__restrict__ probs) {
    const YawRot R = rot[blockIdx.x];
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < n; i += YAW_THREADS) {
        const float4 p = __ldg(&scan[i]);
        // pcl::transformPointCloud with the 4x4 built from the rotation (translation 0), left to right in float
        const float x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R.c, p.x), __fmul_rn(-R.s, p.y)), __fmul_rn(0.f, p.z)), R.tx);
        const float y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R.s, p.x), __fmul_rn(R.c, p.y)), __fmul_rn(0.f, p.z)), R.ty);
        const float z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(0.f, p.x), __fmul_rn(0.f, p.y)), __fmul_rn(R.m22, p.z)), 0.f);
        if (!finite3(x, y, z)) continue;
        const int cx = hmap_cell(x, G.min_x, G.res), cy = hmap_cell(y, G.min_y, G.res);
        if (cx < 0 || cy < 0 || cx >= G.width || cy >= G.height) continue;
        const size_t cell = (size_t)cx * G.height + cy;
        if (__ldg(&cnt[cell]) == 0) continue;
        const double d = (double)__fsub_rn(z, __ldg(&mu[cell]));
        // exp(-pow(z - mu, 2) / (2 * sigma)): sigma == 0 (single-point cell) gives exp(-inf) = 0, or NaN when z == mu
        acc += exp(__ddiv_rn(-__dmul_rn(d, d), (double)__fmul_rn(2.0f, __ldg(&sigma[cell]))));
    }
    __shared__ double part[YAW_THREADS];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = YAW_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) probs[blockIdx.x] = part[0];
}

}  // namespace b2

using namespace b2;

struct b2hmap {
    int device = 0;
    double res = 0.8;
    cudaStream_t st = nullptr;
    VoxPipeline pipe;
    HmapGrid G;
    bool built = false;
    float min_xyz[3], max_xyz[3];
    DevBuf mu, sigma, cnt, misc, rot, probs;
    PinBuf h_small;
};

extern "C" int b2hmap_create(int device, double grid_resolution, b2hmap **out) {
    if (!out) { set_error("b2hmap_create: out is NULL"); return B2_ERR_INVALID; }
    *out = nullptr;
    if (!(grid_resolution > 0)) { set_error("b2hmap_create: resolution must be > 0"); return B2_ERR_INVALID; }
    int rc = check_device(device);
    if (rc) return rc;
    B2_CUDA(cudaSetDevice(device));
    b2hmap *h = new b2hmap();
    h->device = device;
    h->res = grid_resolution < 0.1 ? 0.1 : grid_resolution;      // matching.cpp:138
    cudaError_t e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e)); delete h; return B2_ERR_CUDA; }
    memset(&h->G, 0, sizeof(h->G));
    *out = h;
    return 0;
}

extern "C" void b2hmap_destroy(b2hmap *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->st);
    h->pipe.release();
    h->mu.release(); h->sigma.release(); h->cnt.release(); h->misc.release(); h->rot.release(); h->probs.release();
    h->h_small.release();
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

extern "C" int b2hmap_build(b2hmap *h, b2cloud *local_map, const float origin[3]) {
    if (!h || !local_map || !origin) { set_error("b2hmap_build: NULL argument"); return B2_ERR_INVALID; }
    if (local_map->device != h->device) { set_error("b2hmap_build: handle and cloud live on different devices"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(h->device));
    h->built = false;
    const size_t n = local_map->n;
    int rc;
    if ((rc = h->misc.reserve(256))) return rc;
    if ((rc = h->h_small.reserve(256))) return rc;
    uint32_t *hm = h->h_small.as<uint32_t>();
    for (int a = 0; a < 3; ++a) { hm[a] = f2ord(FLT_MAX); hm[3 + a] = f2ord(-FLT_MAX); }      // SetInitPose: +-max
    hm[6] = 0;
    B2_CUDA(cudaMemcpyAsync(h->misc.p, hm, 32, cudaMemcpyHostToDevice, h->st));
    if (n) {
        unsigned blocks = (unsigned)((n + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        hmap_bounds_kernel<<<blocks, 256, 0, h->st>>>(local_map->d(), (uint32_t)n, origin[0], origin[1], origin[2], h->misc.as<uint32_t>());
        B2_LAUNCH_CHECK();
    }
    B2_CUDA(cudaMemcpyAsync(hm + 8, h->misc.p, 24, cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaStreamSynchronize(h->st));
    for (int a = 0; a < 3; ++a) { h->min_xyz[a] = ord2f(hm[8 + a]); h->max_xyz[a] = ord2f(hm[8 + 3 + a]); }
    // local_map_width_ = std::round((max - min) / res): float difference, double division (matching.cpp:357-358)
    const double w = std::round((double)(h->max_xyz[0] - h->min_xyz[0]) / h->res);
    const double hh = std::round((double)(h->max_xyz[1] - h->min_xyz[1]) / h->res);
    HmapGrid G;
    G.min_x = h->min_xyz[0]; G.min_y = h->min_xyz[1]; G.res = h->res;
    G.width = (n && w > 0 && w < 65536.0) ? (int32_t)w : 0;
    G.height = (n && hh > 0 && hh < 65536.0) ? (int32_t)hh : 0;
    if (n && (w >= 65536.0 || hh >= 65536.0)) { set_error("b2hmap_build: grid %g x %g cells is too large", w, hh); return B2_ERR_INVALID; }
    h->G = G;
    const size_t cells = (size_t)G.width * (size_t)G.height;
    if ((rc = h->mu.reserve(cells * 4 + 16))) return rc;
    if ((rc = h->sigma.reserve(cells * 4 + 16))) return rc;
    if ((rc = h->cnt.reserve(cells * 4 + 16))) return rc;
    h->built = true;
    if (!cells) return 0;
    B2_CUDA(cudaMemsetAsync(h->mu.p, 0, cells * 4, h->st));
    B2_CUDA(cudaMemsetAsync(h->sigma.p, 0, cells * 4, h->st));
    B2_CUDA(cudaMemsetAsync(h->cnt.p, 0, cells * 4, h->st));
    uint32_t off[2] = {0u, (uint32_t)n};
    if ((rc = h->pipe.plan(off, 1, h->st))) return rc;
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    uint32_t *n_valid = h->misc.as<uint32_t>() + 6;
    hmap_key_kernel<<<blocks, 256, 0, h->st>>>(local_map->d(), (uint32_t)n, origin[0], origin[1], G, h->pipe.keys0(), n_valid);
    B2_LAUNCH_CHECK();
    int nbits = 1;
    while (nbits < 32 && (1ull << nbits) <= cells) ++nbits;          // keys lie in [0, cells]
    if ((rc = h->pipe.run_prepared((uint32_t)cells, nbits, h->st))) return rc;
    hmap_cell_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->st>>>(local_map->d(), h->pipe.view(), h->pipe.run_start(), n_valid, origin[2],
                                                                     h->mu.as<float>(), h->sigma.as<float>(), h->cnt.as<int32_t>());
    B2_LAUNCH_CHECK();
    B2_CUDA(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int b2hmap_info(b2hmap *h, int32_t *width, int32_t *height, float min_xyz[3], float max_xyz[3]) {
    if (!h) { set_error("b2hmap_info: NULL handle"); return B2_ERR_INVALID; }
    if (!h->built) { set_error("b2hmap_info: no map built"); return B2_ERR_STATE; }
    if (width) *width = h->G.width;
    if (height) *height = h->G.height;
    for (int a = 0; a < 3; ++a) { if (min_xyz) min_xyz[a] = h->min_xyz[a]; if (max_xyz) max_xyz[a] = h->max_xyz[a]; }
    return 0;
}

extern "C" int b2hmap_cells(b2hmap *h, float *mu, float *sigma, int32_t *point_cnt) {
    if (!h) { set_error("b2hmap_cells: NULL handle"); return B2_ERR_INVALID; }
    if (!h->built) { set_error("b2hmap_cells: no map built"); return B2_ERR_STATE; }
    const size_t cells = (size_t)h->G.width * (size_t)h->G.height;
    if (!cells) return 0;
    B2_CUDA(cudaSetDevice(h->device));
    B2_CUDA(cudaStreamSynchronize(h->st));
    if (mu) B2_CUDA(cudaMemcpy(mu, h->mu.p, cells * 4, cudaMemcpyDeviceToHost));
    if (sigma) B2_CUDA(cudaMemcpy(sigma, h->sigma.p, cells * 4, cudaMemcpyDeviceToHost));
    if (point_cnt) B2_CUDA(cudaMemcpy(point_cnt, h->cnt.p, cells * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// (x, y, yaw) hypothesis scoring: for every offset (dx, dy) of the sensor position relative to the grid origin and
// every yaw bin, the score getInitialYawAngle computes for that pose.  probs[o * angle_size + bin].
extern "C" int b2hmap_pose_search(b2hmap *h, b2cloud *scan, int angle_size, const float *offsets_xy, int n_offsets, double *probs) {
    if (!h || !scan || !probs || (n_offsets > 0 && !offsets_xy)) { set_error("b2hmap_pose_search: NULL argument"); return B2_ERR_INVALID; }
    if (!h->built) { set_error("b2hmap_pose_search: no map built"); return B2_ERR_STATE; }
    if (angle_size < 1 || angle_size > 65535 || n_offsets < 1 || (long long)angle_size * n_offsets > (1 << 24)) {
        set_error("b2hmap_pose_search: angle_size / n_offsets out of range"); return B2_ERR_INVALID;
    }
    if (scan->device != h->device) { set_error("b2hmap_pose_search: handle and cloud live on different devices"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(h->device));
    const size_t nh = (size_t)angle_size * (size_t)n_offsets;
    int rc;
    if ((rc = h->rot.reserve(nh * sizeof(YawRot)))) return rc;
    if ((rc = h->probs.reserve(nh * 8))) return rc;
    const size_t rot_bytes = ((nh * sizeof(YawRot) + 63) / 64) * 64;
    if ((rc = h->h_small.reserve(rot_bytes + nh * 8 + 64))) return rc;
    YawRot *hr = h->h_small.as<YawRot>();
    const float delta_angle = (float)(2 * M_PI / angle_size);          // float delta_angle = 2 * M_PI / angle_size
    for (int o = 0; o < n_offsets; ++o)
        for (int i = 0; i < angle_size; ++i) {
            // Eigen::AngleAxisf(delta_angle * i, UnitZ).matrix() (AngleAxis.h toRotationMatrix): c, s in float,
            // diagonal = (1 - c) * axis^2 + c; float sin / cos evaluated as float(f(double(a))): correctly rounded in
            // practice, platform independent
            const float a = delta_angle * (float)i;
            const float s = (float)std::sin((double)a), c = (float)std::cos((double)a);
            YawRot &r = hr[(size_t)o * angle_size + i];
            r.c = c; r.s = s; r.m22 = (1.0f - c) * 1.0f * 1.0f + c;
            r.tx = offsets_xy[2 * o]; r.ty = offsets_xy[2 * o + 1];
        }
    B2_CUDA(cudaMemcpyAsync(h->rot.p, hr, nh * sizeof(YawRot), cudaMemcpyHostToDevice, h->st));
    const size_t cells = (size_t)h->G.width * (size_t)h->G.height;
    for (size_t i = 0; i < nh; ++i) probs[i] = 0.0;
    if (cells && scan->n) {
        hmap_yaw_kernel<<<(unsigned)nh, YAW_THREADS, 0, h->st>>>(scan->d(), (uint32_t)scan->n, h->G, h->mu.as<float>(), h->sigma.as<float>(),
                                                                h->cnt.as<int32_t>(), h->rot.as<YawRot>(), h->probs.as<double>());
        B2_LAUNCH_CHECK();
        double *hp = reinterpret_cast<double *>(h->h_small.as<char>() + rot_bytes);
        B2_CUDA(cudaMemcpyAsync(hp, h->probs.p, nh * 8, cudaMemcpyDeviceToHost, h->st));
        B2_CUDA(cudaStreamSynchronize(h->st));
        for (size_t i = 0; i < nh; ++i) probs[i] = hp[i];
    } else {
        B2_CUDA(cudaStreamSynchronize(h->st));
    }
    return 0;
}

extern "C" int b2hmap_yaw_search(b2hmap *h, b2cloud *scan, int angle_size, double *probs, double *best_angle) {
    if (!h || !scan || !best_angle) { set_error("b2hmap_yaw_search: NULL argument"); return B2_ERR_INVALID; }
    if (angle_size < 1 || angle_size > 65535) { set_error("b2hmap_yaw_search: angle_size out of range"); return B2_ERR_INVALID; }
    std::vector<double> pr((size_t)angle_size, 0.0);
    const float zero[2] = {0.f, 0.f};
    int rc = b2hmap_pose_search(h, scan, angle_size, zero, 1, pr.data());
    if (rc) return rc;
    const float delta_angle = (float)(2 * M_PI / angle_size);
    // first strictly larger bin wins; NaN bins never win (matching.cpp:298-306)
    float max_prob = -FLT_MAX;
    double best = 0.0;
    for (int it = 0; it < angle_size; ++it)
        if (pr[it] > max_prob) { max_prob = (float)pr[it]; best = it * delta_angle; }
    *best_angle = best;
    if (probs) for (int i = 0; i < angle_size; ++i) probs[i] = pr[i];
    return 0;
}
