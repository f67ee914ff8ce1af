// b2_ndt.cu -- NDT registration on sm_100a: target-grid build (VoxelGridCovariance), the fused
// transform + neighbour gather + score/gradient/Hessian kernel with the on-device Newton /
// More-Thuente loop, the fitness-score kernel, and the b2ndt_* C ABI.
//
// Reference semantics (paths relative to /root/reference/lidar_localization/):
//   NDTRegistration::{SetInputTarget, ScanMatch, GetFitnessScore}  src/models/registration/ndt_registration.cpp:46-66
//   -> pcl::NormalDistributionsTransform 1.7 (restated in-tree at
//      src/models/registration/ndt_registration_manual/NormalDistributionsTransform.cpp [NDTM] and VoxelGrid.cpp).
//
// Data layout in HBM (one target):
//   pts_sorted  float4[N]      target points permuted into voxel order (fitness NN buckets)
//   leaf_idx    int32[V]       voxel linear index, ascending;  leaf_n int32[V];  leaf_start uint32[V]
//   centroid4   float4[V]      float centroid (the kd-tree search point of PCL), intensity mean in .w
//   gauss       double[12][V]  96-byte records {mean[3], icov xx | xy,xz,yy,yz | zz, pad[3]}: two 256-bit loads + one 64-bit
//   cells       float4[ncells] dense grid record {centroid x,y,z, code}: code = +(leaf+1) searchable (n >= min_pts),
//                              -(leaf+1) sparse, 0 empty
//   nbr_head    uint2[ncells]  {first entry, count} of the cell's neighbour list
//   nbr_list    float4[..]     per cell, the searchable leaves of its 3x3x3 window {centroid x,y,z, leaf} in
//                              fixed (z,y,x) order, one contiguous run per cell (a radius query reads ONE
//                              header and one contiguous run instead of 27 scattered cells)
// The path is a gather + reduction (no dense contraction): no tensor cores by design.
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "b2_common.cuh"
#include "b2_ndt_math.cuh"
#include "b2_voxel.cuh"
#include "b2_cloud.cuh"

namespace cg = cooperative_groups;

namespace b2 {
int check_device(int device);

struct TargetDev {
    VoxLayout L;              // host copy of the layout
    uint32_t N = 0, V = 0, n_tree = 0;
    float max_disp = 0.f;     // max distance of a searchable leaf's float centroid outside its own cell
    DevBuf pts_in;            // float4[N] as given (host path)
    DevBuf pts_sorted;        // float4[N]
    DevBuf leaf_idx, leaf_n, leaf_start, centroid4, gauss, sums, icov9, cells, counters;
    DevBuf nbr_head, nbr_list, nbr_tiles;
    bool valid = false;
    // incremental updates (b2ndt_update_target): the float centroid SUMS per leaf, the run -> leaf table of the last
    // update, and the points in the order they were handed over (retained pointer until the first update, like PCL's
    // target_ pointer; an owned copy from then on)
    DevBuf csum4, touched, pts_all;
    const float4 *src = nullptr;
    size_t n_all = 0;
    bool owns_all = false;
    bool stale_buckets = false;   // pts_sorted / leaf_start no longer cover every point (fitness rebuilds first)
    bool unsorted = false;        // leaves appended by updates: the table is no longer in ascending voxel order
    uint32_t upd_incremental = 0, upd_rebuilt = 0;
    // what the dense arrays hold: cells[] is non-zero exactly at leaf_idx[0 .. fp_V), nbr_head[] exactly at the first
    // fp_active entries of the listed-cell table (nbr_tiles); fp_clean = that statement holds (false after a failed build)
    uint32_t fp_V = 0, fp_active = 0;
    bool fp_clean = false;
};

// ------------------------------------------------------------------ target build kernels -----
// Per occupied voxel, PCL's per-leaf accumulation in PCL's order and precision: mean_ += p (double),
// cov_ += p p^T (double, from Identity), centroid += (x,y,z,i) (float), members in input order (the stable sort
// keeps it).  Two kernels, like the VoxelFilter centroids (b2_voxel.cu):
//   leaf_stats_kernel   a warp takes 32 consecutive voxels; a voxel of up to LS_SEQ points is accumulated by ONE
//                       lane; crowded voxels go to a work list;
//   leaf_crowded_kernel one warp per crowded voxel: lanes gather 64 members at a time one group ahead of the fold,
//                       every lane folds them sequentially through shuffles.
constexpr uint32_t LS_SEQ = 32;

struct LeafOut {
    float4 *pts_sorted;
    int32_t *leaf_idx, *leaf_n;
    uint32_t *leaf_start;
    float4 *centroid4;
    double *sums;
    float4 *csum4;       // float sums of x, y, z, intensity (an update continues them)
};

// Incremental update (VoxelGrid::update, reference VoxelGrid.cpp:545-584,736-809): a run of the NEW cloud continues the
// sums of the leaf that already owns its cell, or opens a new leaf behind the existing ones.
struct LeafUpd {
    const float4 *cells;     // dense grid record: .w = +-(leaf + 1), 0 = empty
    const int32_t *leaf_n;   // current point counts
    uint32_t V_old;
    uint32_t *new_ctr;       // leaves opened by this update
    uint32_t *touched;       // run -> leaf
};

__device__ __forceinline__ void leaf_emit(const LeafOut &O, uint32_t j, uint32_t s, uint32_t n, uint32_t key, float cx, float cy,
                                          float cz, float ci, double sx, double sy, double sz, double cxx, double cxy, double cxz,
                                          double cyy, double cyz, double czz) {
    const float fn = (float)n;
    O.leaf_idx[j] = (int32_t)key;
    O.leaf_n[j] = (int32_t)n;
    O.leaf_start[j] = s;
    O.centroid4[j] = make_float4(__fdiv_rn(cx, fn), __fdiv_rn(cy, fn), __fdiv_rn(cz, fn), __fdiv_rn(ci, fn));
    O.csum4[j] = make_float4(cx, cy, cz, ci);
    double *o = O.sums + (size_t)j * 9;
    o[0] = sx; o[1] = sy; o[2] = sz; o[3] = cxx; o[4] = cxy; o[5] = cxz; o[6] = cyy; o[7] = cyz; o[8] = czz;
}

// update mode: the leaf a run of new points belongs to (lane-uniform callers pass the run's key); n0 = its points so far
__device__ __forceinline__ uint32_t upd_find_leaf(const LeafUpd &U, uint32_t key, uint32_t run, uint32_t &n0) {
    const int code = __float_as_int(__ldg(&U.cells[key]).w);
    uint32_t j;
    if (code != 0) { j = (uint32_t)((code > 0 ? code : -code) - 1); n0 = (uint32_t)U.leaf_n[j]; }
    else { j = U.V_old + atomicAdd(U.new_ctr, 1u); n0 = 0u; }
    U.touched[run] = j;
    return j;
}

#define B2_LEAF_ACC(x, y, z, w)                                                                        \
    do {                                                                                               \
        cx = __fadd_rn(cx, x); cy = __fadd_rn(cy, y); cz = __fadd_rn(cz, z); ci = __fadd_rn(ci, w);    \
        const double dx = (double)(x), dy = (double)(y), dz = (double)(z);                             \
        sx = __dadd_rn(sx, dx); sy = __dadd_rn(sy, dy); sz = __dadd_rn(sz, dz);                        \
        cxx = __dadd_rn(cxx, __dmul_rn(dx, dx)); cxy = __dadd_rn(cxy, __dmul_rn(dx, dy));              \
        cxz = __dadd_rn(cxz, __dmul_rn(dx, dz)); cyy = __dadd_rn(cyy, __dmul_rn(dy, dy));              \
        cyz = __dadd_rn(cyz, __dmul_rn(dy, dz)); czz = __dadd_rn(czz, __dmul_rn(dz, dz));              \
    } while (0)

template <bool UPD>
__global__ void __launch_bounds__(256) leaf_stats_kernel(const float4 *__restrict__ pts, SortView sv,
                                                         const uint32_t *__restrict__ run_start, uint32_t V, uint32_t n_finite,
                                                         LeafOut O, uint32_t *__restrict__ crowded, uint32_t *__restrict__ n_crowded,
                                                         LeafUpd U) {
    const uint32_t *__restrict__ keys = sv.keys();
    const uint32_t *__restrict__ vals = sv.vals();
    const int l = threadIdx.x & 31;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t j0 = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32u; j0 < V; j0 += nwarps * 32u) {
        const uint32_t j = j0 + l;
        const bool valid = j < V;
        uint32_t s = 0, e = 0;
        if (valid) { s = run_start[j]; e = (j + 1 < V) ? run_start[j + 1] : n_finite; }
        const uint32_t len = e - s;
        const bool is_crowded = len > LS_SEQ;
        {
            const uint32_t m = __ballot_sync(0xffffffffu, is_crowded);
            if (m) {
                uint32_t base = 0;
                if (l == 0) base = atomicAdd(n_crowded, (uint32_t)__popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (is_crowded) crowded[base + __popc(m & ((1u << l) - 1u))] = j;
            }
        }
        if (valid && !is_crowded) {
            float cx = 0.f, cy = 0.f, cz = 0.f, ci = 0.f;
            double sx = 0, sy = 0, sz = 0, cxx = 1, cxy = 0, cxz = 0, cyy = 1, cyz = 0, czz = 1;
            uint32_t jo = j, n0 = 0;
            if (UPD) {
                jo = upd_find_leaf(U, keys[s], j, n0);
                if (n0) {       // continue the leaf's sums where the earlier clouds left them
                    const float4 c = O.csum4[jo];
                    const double *o = O.sums + (size_t)jo * 9;
                    cx = c.x; cy = c.y; cz = c.z; ci = c.w;
                    sx = o[0]; sy = o[1]; sz = o[2]; cxx = o[3]; cxy = o[4]; cxz = o[5]; cyy = o[6]; cyz = o[7]; czz = o[8];
                }
            }
            uint32_t k = s;
            for (; k + 4 <= e; k += 4) {
                const uint32_t v0 = vals[k], v1 = vals[k + 1], v2 = vals[k + 2], v3 = vals[k + 3];
                const float4 p0 = __ldg(&pts[v0]), p1 = __ldg(&pts[v1]), p2 = __ldg(&pts[v2]), p3 = __ldg(&pts[v3]);
                if (!UPD) { O.pts_sorted[k] = p0; O.pts_sorted[k + 1] = p1; O.pts_sorted[k + 2] = p2; O.pts_sorted[k + 3] = p3; }
                B2_LEAF_ACC(p0.x, p0.y, p0.z, p0.w);
                B2_LEAF_ACC(p1.x, p1.y, p1.z, p1.w);
                B2_LEAF_ACC(p2.x, p2.y, p2.z, p2.w);
                B2_LEAF_ACC(p3.x, p3.y, p3.z, p3.w);
            }
            for (; k < e; ++k) {
                const float4 p = __ldg(&pts[vals[k]]);
                if (!UPD) O.pts_sorted[k] = p;
                B2_LEAF_ACC(p.x, p.y, p.z, p.w);
            }
            leaf_emit(O, jo, s, n0 + len, keys[s], cx, cy, cz, ci, sx, sy, sz, cxx, cxy, cxz, cyy, cyz, czz);
        }
    }
}

constexpr int LCROWD_THREADS = 128;    // 4 warps, each with a ring of LCROWD_RING staging groups of LCROWD_GROUP points
constexpr int LCROWD_GROUP = 128;
constexpr int LCROWD_RING = 4;
template <bool UPD>
__global__ void __launch_bounds__(LCROWD_THREADS) leaf_crowded_kernel(const float4 *__restrict__ pts, SortView sv,
                                                                      const uint32_t *__restrict__ run_start, uint32_t V,
                                                                      uint32_t n_finite, LeafOut O,
                                                                      const uint32_t *__restrict__ crowded,
                                                                      const uint32_t *__restrict__ n_crowded, LeafUpd U) {
    chain_sync();
    const uint32_t n = *n_crowded;
    const uint32_t *__restrict__ keys = sv.keys();
    const uint32_t *__restrict__ vals = sv.vals();
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    // The 13 sums of a leaf are spread over the lanes: lane a < 9 owns one double sum  acc += u * v  with (u, v) =
    // (x,1) (y,1) (z,1) (x,x) (x,y) (x,z) (y,y) (y,z) (z,z)  [mean_ += p and cov_ += p p^T; a product by 1.0 is exact],
    // lanes 9..12 own the float centroid sums of x, y, z, intensity.  A member then costs two 4-byte shared loads, one
    // DMUL, one DADD and one FADD per lane instead of 15 FP64 operations executed redundantly by every lane; every sum
    // is still formed in input order.  The members reach shared memory through asynchronous copies, three groups
    // ahead of the fold.
    __shared__ float4 stage[LCROWD_THREADS / 32][LCROWD_RING][LCROWD_GROUP];
    __shared__ float s_one;
    if (threadIdx.x == 0) s_one = 1.0f;
    __syncthreads();
    // (u, v) component offsets in floats; v == -1: the constant 1
    const int a = (l < 13) ? l : 12;
    const int iu = (a == 0 || a == 3 || a == 4 || a == 5 || a == 9) ? 0 : (a == 1 || a == 6 || a == 7 || a == 10) ? 1 : (a == 12) ? 3 : 2;
    const int iv = (a == 3) ? 0 : (a == 4 || a == 6) ? 1 : (a == 5 || a == 7 || a == 8) ? 2 : -1;
    for (uint32_t i = w * gridDim.x + blockIdx.x; i < n; i += nwarps) {
        const uint32_t j = crowded[i];
        const uint32_t s = run_start[j];
        const uint32_t e = (j + 1 < V) ? run_start[j + 1] : n_finite;
        auto ldv = [&](uint32_t c, uint32_t (&v)[4]) {
#pragma unroll
            for (int d = 0; d < 4; ++d) { const uint32_t idx = c + d * 32 + l; v[d] = (idx < e) ? __ldg(&vals[idx]) : 0xFFFFFFFFu; }
        };
        auto gather = [&](const uint32_t (&v)[4], int buf) {
#pragma unroll
            for (int d = 0; d < 4; ++d)
                if (v[d] != 0xFFFFFFFFu) cp_async16(&stage[w][buf][d * 32 + l], &pts[v[d]]);
            cp_async_commit();
        };
        double dacc = (a == 3 || a == 6 || a == 8) ? 1.0 : 0.0;       // cov_ starts from Identity
        float facc = 0.f;
        uint32_t jo = j, n0 = 0;
        if (UPD) {
            if (l == 0) jo = upd_find_leaf(U, keys[s], j, n0);
            jo = __shfl_sync(0xffffffffu, jo, 0); n0 = __shfl_sync(0xffffffffu, n0, 0);
            if (n0) {
                if (l < 9) dacc = O.sums[(size_t)jo * 9 + l];
                else if (l < 13) facc = reinterpret_cast<const float *>(&O.csum4[jo])[l - 9];
            }
        }
        auto fold = [&](int buf, uint32_t c) {
            const int m = (e - c < (uint32_t)LCROWD_GROUP) ? (int)(e - c) : LCROWD_GROUP;
            // the target keeps its points in voxel order (fitness buckets): this group's slice, coalesced
            if (!UPD) {
#pragma unroll
                for (int d = 0; d < 4; ++d)
                    if (d * 32 + l < m) O.pts_sorted[c + d * 32 + l] = stage[w][buf][d * 32 + l];
            }
            const float *qu = reinterpret_cast<const float *>(stage[w][buf]) + iu;
            const float *qv = (iv >= 0) ? reinterpret_cast<const float *>(stage[w][buf]) + iv : &s_one;
            const int v_stride = (iv >= 0) ? 4 : 0;
            int k = 0;
            for (; k + 16 <= m; k += 16) {
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const float u = qu[4 * (k + t)], v = qv[v_stride * (k + t)];
                    facc = __fadd_rn(facc, u);
                    dacc = __dadd_rn(dacc, __dmul_rn((double)u, (double)v));
                }
            }
#pragma unroll 4
            for (; k < m; ++k) {
                const float u = qu[4 * k], v = qv[v_stride * k];
                facc = __fadd_rn(facc, u);
                dacc = __dadd_rn(dacc, __dmul_rn((double)u, (double)v));
            }
        };
        uint32_t v0[4], v1[4], v2[4], vn[4];
        ldv(s, v0); ldv(s + LCROWD_GROUP, v1); ldv(s + 2 * LCROWD_GROUP, v2); ldv(s + 3 * LCROWD_GROUP, vn);
        gather(v0, 0); gather(v1, 1); gather(v2, 2);
        int buf = 0;
        for (uint32_t c = s; c < e; c += LCROWD_GROUP) {
            gather(vn, (buf + 3) & (LCROWD_RING - 1));
            ldv(c + 4 * LCROWD_GROUP, vn);
            cp_async_wait<3>();
            __syncwarp();
            fold(buf, c);
            __syncwarp();
            buf = (buf + 1) & (LCROWD_RING - 1);
        }
        cp_async_wait<0>();
        double d9[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) d9[k] = __shfl_sync(0xffffffffu, dacc, k);
        const float cx = __shfl_sync(0xffffffffu, facc, 9), cy = __shfl_sync(0xffffffffu, facc, 10);
        const float cz = __shfl_sync(0xffffffffu, facc, 11), ci = __shfl_sync(0xffffffffu, facc, 12);
        if (l == 0) leaf_emit(O, jo, s, n0 + (e - s), keys[s], cx, cy, cz, ci, d9[0], d9[1], d9[2], d9[3], d9[4], d9[5], d9[6], d9[7], d9[8]);
    }
}

// One thread per voxel: mean, single-pass covariance, eigen inflation, inverse (leaf_finish), the
// 80-byte gather record, and the dense-grid entry.
struct LayoutArg { int32_t min_b[3], div_b[3]; float inv[3]; float res; };
constexpr int GAUSS_STRIDE = 12;     // doubles per voxel record (96 bytes, 32-byte aligned)

// Can the leaf with float centroid (cx,cy,cz) be within the search radius of ANY query that falls into
// cell (kx,ky,kz)?  Distance from the centroid to the cell's box, against the radius plus a generous slack
// (covers the float rounding of the query's cell assignment and of the distance test).  Prunes ~1/4 of
// the 3x3x3 window (corner and edge cells); used identically by the count and the fill pass.
__device__ __forceinline__ bool nbr_keep(float cx, float cy, float cz, int kx, int ky, int kz, const LayoutArg &LA) {
    const double c[3] = {(double)cx, (double)cy, (double)cz};
    const int k[3] = {kx, ky, kz};
    double d2 = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double lo = (double)(k[a] + LA.min_b[a]) / (double)LA.inv[a];
        const double hi = (double)(k[a] + LA.min_b[a] + 1) / (double)LA.inv[a];
        const double d = fmax(fmax(lo - c[a], c[a] - hi), 0.0);
        d2 += d * d;
    }
    const double lim = (double)LA.res * 1.002 + 1e-5 * (fabs(c[0]) + fabs(c[1]) + fabs(c[2]));
    return d2 < lim * lim;
}

__device__ __forceinline__ void leaf_nbr_account(uint32_t j, const LayoutArg &LA, const int32_t *__restrict__ leaf_idx,
                                                 const float4 *__restrict__ centroid4, uint2 *__restrict__ nbr_head,
                                                 uint32_t *__restrict__ counters, uint32_t *__restrict__ active);

__global__ void __launch_bounds__(128) leaf_finish_kernel(uint32_t V, int min_pts, double eig_mult, LayoutArg LA,
                                                          const int32_t *__restrict__ leaf_idx, const int32_t *__restrict__ leaf_n,
                                                          const float4 *__restrict__ centroid4,
                                                          const double *__restrict__ sums, double *__restrict__ gauss,
                                                          double *__restrict__ icov9, float4 *__restrict__ cells,
                                                          uint2 *__restrict__ nbr_head, uint32_t *__restrict__ counters,
                                                          uint32_t *__restrict__ active, const uint32_t *__restrict__ list) {
    // list != NULL (incremental update): V entries of `list` name the leaves to finish, and the neighbour-list
    // accounting is left to nbr_count_kernel (it has to be redone over ALL leaves)
    chain_sync();
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= V) return;
    const bool account = (list == nullptr);
    if (list) j = list[j];
    const double *s = sums + (size_t)j * 9;
    double sum[3] = {s[0], s[1], s[2]};
    double acc[9] = {s[3], s[4], s[5], s[4], s[6], s[7], s[5], s[7], s[8]};
    double mean[3], cov[9], icov[9], ev[3];
    const int n = leaf_n[j];
    // a leaf that fails PCL's eigenvalue checks (nr_points = -1) stays in the centroid kd-tree with the
    // icov_ it has at that moment (zero, or the non-finite inverse), exactly what leaf_finish leaves here
    (void)leaf_finish(sum, acc, n, min_pts, eig_mult, mean, cov, icov, ev);
    double *g = gauss + (size_t)j * GAUSS_STRIDE;
    g[0] = mean[0]; g[1] = mean[1]; g[2] = mean[2];
    g[3] = icov[0]; g[4] = icov[1]; g[5] = icov[2]; g[6] = icov[4]; g[7] = icov[5]; g[8] = icov[8];
    g[9] = 0.0; g[10] = 0.0; g[11] = 0.0;
    double *ic = icov9 + (size_t)j * 9;
    for (int a = 0; a < 9; ++a) ic[a] = icov[a];
    const bool tree = n >= min_pts;
    {
        // dense-grid record: the float centroid travels with the voxel code so the search needs one load
        const float4 c = centroid4[j];
        cells[leaf_idx[j]] = make_float4(c.x, c.y, c.z, __int_as_float(tree ? (int32_t)(j + 1) : -(int32_t)(j + 1)));
    }
    if (tree && account) leaf_nbr_account(j, LA, leaf_idx, centroid4, nbr_head, counters, active);
}

// searchable leaf j: count it, add it to the list length of every cell it can be reached from, and fold how far its
// float centroid lies outside its own cell into the maximum
__device__ __forceinline__ void leaf_nbr_account(uint32_t j, const LayoutArg &LA, const int32_t *__restrict__ leaf_idx,
                                                 const float4 *__restrict__ centroid4, uint2 *__restrict__ nbr_head,
                                                 uint32_t *__restrict__ counters, uint32_t *__restrict__ active) {
    {
        uint32_t first_mask = 0u;
        // how far the float centroid (PCL's kd-tree point) lies outside its own cell: bounds the search
        // window margin of the match kernel.  Cell k of an axis spans [k/inv, (k+1)/inv).
        const int idx = leaf_idx[j];
        const int iz = idx / (LA.div_b[0] * LA.div_b[1]);
        const int iy = (idx - iz * LA.div_b[0] * LA.div_b[1]) / LA.div_b[0];
        const int ix = idx - iz * LA.div_b[0] * LA.div_b[1] - iy * LA.div_b[0];
        // this leaf appears in the neighbour list of every in-grid cell of its 3x3x3 window it can be reached from
        const float4 c0 = centroid4[j];
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int kx = ix + dx, ky = iy + dy, kz = iz + dz;
                    if (kx < 0 || ky < 0 || kz < 0 || kx >= LA.div_b[0] || ky >= LA.div_b[1] || kz >= LA.div_b[2]) continue;
                    if (!nbr_keep(c0.x, c0.y, c0.z, kx, ky, kz, LA)) continue;
                    const size_t cell = (size_t)kx + (size_t)ky * LA.div_b[0] + (size_t)kz * LA.div_b[0] * LA.div_b[1];
                    // the first leaf to reach a cell puts it on the list of cells that own a neighbour list (below)
                    if (atomicAdd(&nbr_head[cell].y, 1u) == 0u) first_mask |= 1u << ((dz + 1) * 9 + (dy + 1) * 3 + (dx + 1));
                }
        {
            // one atomicAdd per warp for the listed-cell slots (a counter bumped once per cell serialises the kernel)
            cg::coalesced_group g = cg::coalesced_threads();
            const uint32_t mine = (uint32_t)__popc(first_mask);
            const uint32_t incl = cg::inclusive_scan(g, mine);
            uint32_t base = 0;
            if (g.thread_rank() == g.size() - 1 && incl) base = atomicAdd(&counters[3], incl);
            base = g.shfl(base, g.size() - 1);
            uint32_t o = base + incl - mine;
            for (uint32_t m = first_mask; m; m &= m - 1u) {
                const int b = __ffs(m) - 1;
                const int kz = iz + b / 9 - 1, ky = iy + (b % 9) / 3 - 1, kx = ix + b % 3 - 1;
                active[o++] = (uint32_t)((size_t)kx + (size_t)ky * LA.div_b[0] + (size_t)kz * LA.div_b[0] * LA.div_b[1]);
            }
            if (g.thread_rank() == 0) atomicAdd(&counters[0], g.size());
        }
        const float4 c = centroid4[j];
        const double cc[3] = {(double)c.x, (double)c.y, (double)c.z};
        const int ii[3] = {ix, iy, iz};
        double disp = 0.0;
        for (int a = 0; a < 3; ++a) {
            const double lo = (double)(ii[a] + LA.min_b[a]) / (double)LA.inv[a];
            const double hi = (double)(ii[a] + LA.min_b[a] + 1) / (double)LA.inv[a];
            disp = fmax(disp, fmax(lo - cc[a], cc[a] - hi));
        }
        if (disp > 0.0) atomicMax(&counters[1], __float_as_uint(__double2float_ru(disp)));   // positive floats order as uints
    }
}

// incremental update: the neighbour-list accounting of leaf_finish_kernel redone over every leaf (the list heads were
// cleared: a leaf that became searchable, or whose centroid moved, changes the lists of up to 27 cells)
__global__ void __launch_bounds__(128) nbr_count_kernel(uint32_t V, int min_pts, LayoutArg LA, const int32_t *__restrict__ leaf_idx,
                                                        const int32_t *__restrict__ leaf_n, const float4 *__restrict__ centroid4,
                                                        uint2 *__restrict__ nbr_head, uint32_t *__restrict__ counters,
                                                        uint32_t *__restrict__ active) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= V) return;
    if (leaf_n[j] >= min_pts) leaf_nbr_account(j, LA, leaf_idx, centroid4, nbr_head, counters, active);
}

// incremental update: voxel key of every new point under the EXISTING layout; a finite point outside the grid's index
// box raises *oob (the layout of old + new points differs: the caller rebuilds); non-finite points get the invalid key
__global__ void __launch_bounds__(256) upd_key_kernel(const float4 *__restrict__ pts, uint32_t n, VoxLayout L,
                                                      uint32_t *__restrict__ keys, uint32_t *__restrict__ n_valid,
                                                      uint32_t *__restrict__ oob) {
    uint32_t cnt = 0, bad = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = __ldg(&pts[i]);
        uint32_t key = L.ncells;
        if (finite3(p.x, p.y, p.z)) {
            const int i0 = (int)(floorf(__fmul_rn(p.x, L.inv[0])) - (float)L.min_b[0]);
            const int i1 = (int)(floorf(__fmul_rn(p.y, L.inv[1])) - (float)L.min_b[1]);
            const int i2 = (int)(floorf(__fmul_rn(p.z, L.inv[2])) - (float)L.min_b[2]);
            if (i0 < 0 || i1 < 0 || i2 < 0 || i0 >= L.div_b[0] || i1 >= L.div_b[1] || i2 >= L.div_b[2]) bad = 1u;
            else { key = (uint32_t)(i0 * L.mul[0] + i1 * L.mul[1] + i2 * L.mul[2]); ++cnt; }
        }
        keys[i] = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cnt) atomicAdd(n_valid, cnt);
        if (bad) atomicOr(oob, 1u);
    }
}

// ------------------------------------------------------------------ neighbour lists ----------
// nbr_head[c].y holds the number of searchable leaves in the 3x3x3 window of cell c that can be reached from it, and the
// cells with a non-zero count stand on the `active` list (both written by leaf_nbr_account).  nbr_fill_kernel, one
// thread per listed cell, takes the cell's run of the list array (one atomicAdd per warp) and gathers the entries in
// fixed (z, y, x) window order.  WHERE a cell's run lies depends on the scheduling; its content and order do not, so a
// query reads the same candidates in the same order in every build.
// The dense arrays (cells: 16 B, nbr_head: 8 B per grid cell, mostly empty) are never cleared as a whole after their
// allocation: grid_clear_kernel, first thing in the next build, zeroes exactly the entries the previous build wrote
// (its leaves' cells, its listed cells).
constexpr int NBR_TPB = 256;
constexpr uint32_t NB_MASK = 0x3FFFFFFFu;

// counters: [0] searchable leaves, [1] max centroid displacement (float bits), [3] listed cells, [4] list entries in
// total, [5] crowded voxels, [6] leaves opened by an update, [7] / [8] update: valid new points / outside-the-box flag
__global__ void __launch_bounds__(256) grid_clear_kernel(const int32_t *__restrict__ leaf_idx, uint32_t V,
                                                         const uint32_t *__restrict__ active, uint32_t n_active,
                                                         float4 *__restrict__ cells, uint2 *__restrict__ head) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (cells && i < V) cells[leaf_idx[i]] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n_active) head[active[i]] = make_uint2(0u, 0u);
}

__global__ void __launch_bounds__(NBR_TPB) nbr_fill_kernel(uint2 *__restrict__ head, const float4 *__restrict__ cells,
                                                           uint32_t *__restrict__ counters,
                                                           const uint32_t *__restrict__ active, LayoutArg LA,
                                                           float4 *__restrict__ list) {
    chain_sync();
    const uint32_t n = counters[3];
    const int dx_n = LA.div_b[0], dy_n = LA.div_b[1], dz_n = LA.div_b[2];
    const int lane = threadIdx.x & 31;
    for (uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; i0 < n; i0 += gridDim.x * blockDim.x) {   // warp-uniform
        const uint32_t i = i0 + lane;
        const bool on = i < n;
        const uint32_t c = on ? active[i] : 0u;
        const uint32_t cnt = on ? head[c].y : 0u;
        // this warp's run of the list array
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        uint32_t base = 0;
        if (lane == 31) base = atomicAdd(&counters[4], incl);
        base = __shfl_sync(0xffffffffu, base, 31);
        if (!on) continue;
        const uint32_t off = base + incl - cnt;
        if ((size_t)off + cnt > (size_t)NB_MASK) continue;        // over the entry limit: reported by the host (counters[4])
        head[c].x = off;
        const int iz = (int)(c / ((uint32_t)dx_n * dy_n));
        const int rem = (int)(c - (uint32_t)iz * dx_n * dy_n);
        const int iy = rem / dx_n, ix = rem - iy * dx_n;
        uint32_t k = 0;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy) {
                const int ky = iy + dy, kz = iz + dz;
                if (ky < 0 || kz < 0 || ky >= dy_n || kz >= dz_n) continue;
                const float4 *row = cells + (size_t)ky * dx_n + (size_t)kz * dx_n * dy_n;
                float4 e3[3];
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int kx = ix + dx;
                    e3[dx + 1] = (kx >= 0 && kx < dx_n) ? __ldg(&row[kx]) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int code = __float_as_int(e3[dx].w);
                    if (code > 0 && nbr_keep(e3[dx].x, e3[dx].y, e3[dx].z, ix, iy, iz, LA)) {
                        list[off + k] = make_float4(e3[dx].x, e3[dx].y, e3[dx].z, __int_as_float(code - 1));
                        ++k;
                    }
                }
            }
    }
}

// ------------------------------------------------------------------ the NDT match kernels -----
// The whole Newton / More-Thuente loop of a match runs inside a kernel (no host round trip per iteration).  Two
// kernels share the same building blocks (search_pass, compute_drain, controller_step, below):
//   ndt_batch_kernel  batches: persistent CTAs, NDT_SLOTS matches in flight per CTA, work fetched from a counter;
//   ndt_match_kernel  a single ScanMatch / the derivatives-only entry: one thread-block cluster (1..16 CTAs) per match,
//                     in two CTA shapes (<NDT_NCW> = 4 compute + 8 search warps, two CTAs per SM; <NDT_NCW_WIDE> = 8 + 8,
//                     one CTA per SM, for launches with at most one CTA per SM).
// Every CTA is WARP-SPECIALISED (producer / consumer, registers re-balanced with setmaxnreg):
//   search warps  (NDT_NSW, 56 registers): lane = source point.  Float transform, ONE 8-byte neighbour-list header
//     per query, then the warp walks the concatenation of its 32 lists (coalesced 16-byte entries {centroid, leaf},
//     owner by binary search over the scan of list lengths), float L2 tests, hits {source point, leaf} appended to the
//     warp's ring in shared memory at positions given by ballots (deterministic order, no atomics);
//   compute warps (NDT_NCW, 128 registers): lane = (point, voxel) pair.  Each compute warp drains the rings of its
//     search warps in a fixed round-robin, 32 pairs at a time, so every lane is busy: 96-byte record gather (two
//     256-bit loads + one 64-bit), exp, and the pair's contribution in Q/P/M form (ndt_pair) accumulated in NACC = 35
//     FP64 registers.
// The latency-bound gather and the FP64-bound arithmetic thus run concurrently on different warps, and the cheap
// search warps raise the number of resident warps per SM.
#ifndef NDT_NCW
#define NDT_NCW 4
#endif
#ifndef NDT_NSW
#define NDT_NSW 8
#endif
#ifndef NDT_MIN_CTAS
#define NDT_MIN_CTAS 2
#endif
#ifndef NDT_REG_COMPUTE
#define NDT_REG_COMPUTE 128
#endif
#ifndef NDT_REG_SEARCH
#define NDT_REG_SEARCH 56
#endif
#ifndef NDT_LPF
#define NDT_LPF 1      // search warps pull the neighbour-list lines of the next round into L1
#endif
constexpr int NACC = 35;                       // per-lane accumulators of a derivative pass (layout at ndt_pair)
constexpr int NDT_WARPS = NDT_NCW + NDT_NSW;
constexpr int NDT_THREADS = NDT_WARPS * 32;
constexpr int NDT_PPC = NDT_NSW / NDT_NCW;     // producers (search warps) per compute warp
// A single match (and any launch whose CTAs fit the GPU one per SM) runs ndt_match_kernel with NDT_NCW_WIDE compute warps
// per CTA, one per search warp: the drain of a pass -- the longest phase of a single match, ~113 pairs = 4 dependent
// 32-pair chunks per compute warp -- takes half the steps.  Register file: 8 x 128 + 8 x 56 per lane = one CTA per SM.
constexpr int NDT_NCW_WIDE = (NDT_NSW > NDT_NCW) ? NDT_NSW : NDT_NCW;
#ifndef NDT_POLL_ALL
#define NDT_POLL_ALL 1  // ring polls (consumer: next chunk, producer: ring room) run on every lane (same shared-memory words, broadcast)
                        // instead of lane 0 only: the warp stays converged, so the shuffles that follow never take the
                        // divergent-warp path (WARPSYNC.COLLECTIVE per shuffle: 350 of them in the reduction = 41 us)
#endif
#ifndef NDT_RING
#define NDT_RING 512
#endif
#ifndef NDT_DRAIN_SLEEP
#define NDT_DRAIN_SLEEP 128     // ns between two looks of a compute warp at an empty ring
#endif
#ifndef NDT_ROOM_SLEEP
#define NDT_ROOM_SLEEP 256     // ns between two looks at the consumer position while a search warp's ring is full
#endif
#ifndef NDT_MBAR_HINT
#define NDT_MBAR_HINT 1000     // ns: suspend-time hint of the pass-request wait (0: the hardware's default time limit)
#endif
constexpr uint32_t RING = NDT_RING;            // ring entries per search warp (power of two, >= 128 + 32)
static_assert(NDT_NCW % 4 == 0 && NDT_NSW % 4 == 0 && NDT_NSW % NDT_NCW == 0, "warp-group multiples");

struct GridView {
    const float4 *cells;         // dense grid record {cx, cy, cz, int code}: code > 0 searchable leaf+1, < 0 sparse, 0 empty
    const uint2 *nbr_head;       // per cell {first entry, count} of its neighbour list
    const float4 *nbr_list;      // {cx, cy, cz, leaf} of the searchable leaves of the cell's 3x3x3 window
    const double *gauss;
    int32_t min_b[3], div_b[3], mul[3];
    float res, r2;      // search radius = resolution ; r2 = (float)(res*res)
    float inv_leaf;     // 1/res (only used to find candidate cells)
    float margin;       // >= how far a float centroid can lie outside its own cell (measured at build time)
    int32_t ok;
};

struct MatchArgs {
    const float4 *src;           // packed sources
    const uint32_t *offsets;     // B+1, or nullptr => every match uses [0, n_shared)
    uint32_t n_shared;
    const float *guesses;        // B*16
    const double *poses6;        // B*6 (deriv-only mode)
    float *poses_out;            // B*16
    b2ndt_result *results;       // B
    double *acc_out;             // B*ACC_N (deriv-only mode)
    int deriv_only;
    const uint32_t *ready;       // batch kernel, host-streamed sources: ready[0] = number of leading matches whose points
                                 // have arrived in HBM (written by the copy stream between chunk copies), ready[1] = set
                                 // by the kernel when that wait timed out; NULL = all resident
    unsigned long long *timing;  // NDT_TIMING builds only (tools/ sweeps): per-phase SM-cycle totals of the batch kernel
};

#ifdef NDT_TIMING
#define TMB_DECL unsigned long long tmb[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tmb_t = clock64(); (void)tmb_t
#define TMB_LAP(k) do { const long long _n = clock64(); tmb[k] += (unsigned long long)(_n - tmb_t); tmb_t = _n; } while (0)
#define TMB_FLUSH(base, cond) do { if ((cond) && lane == 0 && A.timing) for (int _k = 0; _k < 8; ++_k) atomicAdd(&A.timing[(base) + _k], tmb[_k]); } while (0)
#else
#define TMB_DECL do { } while (0)
#define TMB_LAP(k) do { } while (0)
#define TMB_FLUSH(base, cond) do { } while (0)
#endif


#ifndef NDT_SLOTS_N
#define NDT_SLOTS_N 2
#endif
constexpr int NDT_SLOTS = NDT_SLOTS_N;          // matches in flight per CTA (batch kernel)
static_assert(NDT_SLOTS >= 1 && NDT_SLOTS <= NDT_NCW && NDT_SLOTS <= 8, "the first match of slot s is started by compute warp s");
constexpr uint32_t PASS_RING = 8;               // pass markers kept per search warp (power of two >= NDT_SLOTS)

struct Slot {                      // one match in flight
    Ctl ctl;
    double warp_part[NDT_NCW_WIDE][NACC];
    double cta_part[2][NACC];      // cluster mode: double-buffered per-CTA partial, read by cluster peers over DSMEM
    double raw_total[NACC];        // reduced pair sums (NACC layout)
    double total[ACC_N];           // contracted with the angle tables: what the controller consumes
    double trig_d[6];              // snapped double sin x3, cos x3 of the requested pose
    float  trig_f[6];              // float sin x3, cos x3
    uint32_t first, last;          // source range of the match
    uint32_t match;                // match index
    uint32_t seq;                  // batch kernel: passes requested so far for this slot (monotonic)
    uint32_t dead;                 // batch kernel: no more work for this slot
    int go;
};

struct NdtSmem {
    Slot slot[NDT_SLOTS];
    // producer -> consumer rings (monotonic counters, never reset)
    uint32_t tail[NDT_NSW];        // entries produced by search warp s
    uint32_t head[NDT_NSW];        // entries consumed from search warp s
    uint32_t finished[NDT_NSW];    // passes search warp s has completed
    uint64_t req_bar[NDT_SLOTS];   // batch kernel: one mbarrier per slot, a phase = one pass request (controller arrives)
    uint32_t pass_end[NDT_NSW][PASS_RING]; // ring position at which pass p of search warp s ended (p % PASS_RING): a producer is
                                   // at most NDT_SLOTS passes ahead of its consumer (a slot's next pass is requested
                                   // only after its previous one was drained by every compute warp)
    uint32_t pass_dead[NDT_NSW][PASS_RING];// batch kernel: "pass" p is only the marker that its slot has no more work
    float4 ring[NDT_NSW][RING];    // (source point x, y, z, leaf index): consumers never touch the source cloud
    float4 stage[NDT_NSW][192];    // per search warp, three slots: [lane] source point, [32 + lane] transformed point
    uint32_t cur_slot[NDT_NCW_WIDE];   // per compute warp: the slot whose pass it is draining
};

static_assert(sizeof(NdtSmem) <= (233472 / NDT_MIN_CTAS) - 1024, "NDT_MIN_CTAS CTAs per SM no longer fit the 228 KB of shared memory");
template <int THREADS>
__device__ __forceinline__ void cta_barrier() { asm volatile("bar.sync 0, %0;" ::"n"(THREADS) : "memory"); }
__device__ __forceinline__ uint32_t ld_vol(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void st_vol(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }
#ifndef NDT_FENCE
#define NDT_FENCE 1
#endif
// producer side: make the ring entries written by the warp (ordered by the preceding __syncwarp) visible
// before the new tail
__device__ __forceinline__ void publish_tail(uint32_t *p, uint32_t v) {
#if NDT_FENCE == 1
    __threadfence_block();
    st_vol(p, v);
#elif NDT_FENCE == 2
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
#else
    st_vol(p, v);
#endif
}

__device__ __forceinline__ double dot3v(const double *h, double x, double y, double z) { return x * h[0] + y * h[1] + z * h[2]; }
__device__ __forceinline__ double dot2v(const double *h, double x, double y) { return x * h[0] + y * h[1]; }

// Per-lane accumulators of a derivative pass (NACC doubles).  With M = w (Sigma^-1 - d2 q q^T), q = Sigma^-1 x',
// w = d1 d2 exp(-d2 x'^T q / 2) and J = [I | C(x)] (computePointDerivatives, NDTM:448-482), the sums of
// updateDerivatives (NDTM:485-520) are
//   gradient = sum w J^T q,   Hessian = sum J^T M J + [second-derivative term  w q . H_E(x)]
// and both C(x) and H_E(x) are LINEAR in the source point x, so the gradient's rotational part and the whole
// second-derivative term are contractions of P = sum w q x^T with the angle tables.  A pair therefore only
// accumulates
//   [0] score  [1..3] Q = sum w q  [4..12] P (3x3, row r = q component)  [13..18] H00 = sum M (xx,xy,xz,yy,yz,zz)
//   [19..27] H01 = sum M C (3x3)  [28..33] H11c = sum C^T M C (33,34,35,44,45,55)  [34] pair count
// (~150 FP64 operations per pair instead of ~250, and no angle-table reads for the second-derivative term);
// acc_finish() contracts P with the tables once per pass and emits the ACC_N-vector the controller consumes.
struct PairRec { double mx, my, mz, ixx, ixy, ixz, iyy, iyz, izz; };

// 96-byte record: two 256-bit loads (LDG.E.256, sm_100) + one 64-bit load = 3 L1 requests instead of 5
__device__ __forceinline__ void ndt_pair_load(const double *__restrict__ g, PairRec &r) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.mx), "=d"(r.my), "=d"(r.mz), "=d"(r.ixx) : "l"(g));
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.ixy), "=d"(r.ixz), "=d"(r.iyy), "=d"(r.iyz) : "l"(g + 4));
    r.izz = __ldg(g + 8);
}

// `on` = this lane holds a pair.  A pair that fails the rejection test of NDTM:499-501 (or an idle lane) contributes
// exact zeros instead of branching away, so two pairs per lane form two independent instruction streams.
__device__ __forceinline__ void ndt_pair_math(const bool on, const float px, const float py, const float pz, const PairRec &r,
                                              const float *__restrict__ T, const AngTab &ang, double d1, double d2, bool hess,
                                              double *acc) {
    float tx, ty, tz;
    transform_f32(T, px, py, pz, tx, ty, tz);
    const double mx = r.mx, my = r.my, mz = r.mz, ixx = r.ixx, ixy = r.ixy, ixz = r.ixz, iyy = r.iyy, iyz = r.iyz, izz = r.izz;
    acc[34] += on ? 1.0 : 0.0;
    const double xq = (double)tx - mx, yq = (double)ty - my, zq = (double)tz - mz;
    double q0 = ixx * xq + ixy * yq + ixz * zq;
    double q1 = ixy * xq + iyy * yq + iyz * zq;
    double q2 = ixz * xq + iyz * yq + izz * zq;
    const double m = xq * q0 + yq * q1 + zq * q2;
    double e = exp(-d2 * m / 2);
    const double sinc = -d1 * e;
    e = d2 * e;
    const bool keep = on && !(e > 1 || e < 0 || e != e);      // NDTM:499-501
    const double w = keep ? e * d1 : 0.0;
    if (!keep) { q0 = 0.0; q1 = 0.0; q2 = 0.0; }              // 0 * inf / NaN must not reach the sums
    acc[0] += keep ? sinc : 0.0;
    const double x = (double)px, y = (double)py, z = (double)pz;
    const double wq0 = w * q0, wq1 = w * q1, wq2 = w * q2;
    acc[1] += wq0; acc[2] += wq1; acc[3] += wq2;
    acc[4] += wq0 * x; acc[5] += wq0 * y; acc[6] += wq0 * z;
    acc[7] += wq1 * x; acc[8] += wq1 * y; acc[9] += wq1 * z;
    acc[10] += wq2 * x; acc[11] += wq2 * y; acc[12] += wq2 * z;
    if (!hess) return;
    // M = w Sigma^-1 - d2 (w q) q^T
    const double t0 = -d2 * wq0, t1 = -d2 * wq1, t2 = -d2 * wq2;
    const double mxx = t0 * q0 + w * ixx, mxy = t0 * q1 + w * ixy, mxz = t0 * q2 + w * ixz;
    const double myy = t1 * q1 + w * iyy, myz = t1 * q2 + w * iyz, mzz = t2 * q2 + w * izz;
    acc[13] += mxx; acc[14] += mxy; acc[15] += mxz; acc[16] += myy; acc[17] += myz; acc[18] += mzz;
    // C = [c3 c4 c5], c3 = (0,J0,J1), c4 = (J2,J3,J4), c5 = (J5,J6,J7)  (j_ang_f/g/h have no z component)
    const double J0 = dot3v(ang.j[0], x, y, z), J1 = dot3v(ang.j[1], x, y, z);
    const double J2 = dot3v(ang.j[2], x, y, z), J3 = dot3v(ang.j[3], x, y, z), J4 = dot3v(ang.j[4], x, y, z);
    const double J5 = dot2v(ang.j[5], x, y), J6 = dot2v(ang.j[6], x, y), J7 = dot2v(ang.j[7], x, y);
    // M C
    const double m30 = mxy * J0 + mxz * J1, m31 = myy * J0 + myz * J1, m32 = myz * J0 + mzz * J1;
    const double m40 = mxx * J2 + mxy * J3 + mxz * J4, m41 = mxy * J2 + myy * J3 + myz * J4, m42 = mxz * J2 + myz * J3 + mzz * J4;
    const double m50 = mxx * J5 + mxy * J6 + mxz * J7, m51 = mxy * J5 + myy * J6 + myz * J7, m52 = mxz * J5 + myz * J6 + mzz * J7;
    acc[19] += m30; acc[20] += m40; acc[21] += m50;
    acc[22] += m31; acc[23] += m41; acc[24] += m51;
    acc[25] += m32; acc[26] += m42; acc[27] += m52;
    // C^T M C
    acc[28] += J0 * m31 + J1 * m32;
    acc[29] += J0 * m41 + J1 * m42;
    acc[30] += J0 * m51 + J1 * m52;
    acc[31] += J2 * m40 + J3 * m41 + J4 * m42;
    acc[32] += J2 * m50 + J3 * m51 + J4 * m52;
    acc[33] += J5 * m50 + J6 * m51 + J7 * m52;
}

// Contract the reduced NACC sums with the angle tables of the finished pass -> ACC_N layout of the controller:
// [0] score, [1..6] gradient, [7..27] Hessian upper triangle row-major, [28] pairs.  Lane o (< ACC_N) of one
// warp computes output o = t[base] + sum_k P[row_k] . table[vec_k]; AngTab is 23 consecutive 3-vectors
// (j_ang_a..h then h_ang_a2..f3).
// (global memory, not __constant__: lane o reads entry o, and constant memory serialises a warp's distinct addresses)
__device__ const signed char ACCF_BASE[ACC_N] = {0, 1, 2, 3, -1, -1, -1,
                                             13, 14, 15, 19, 20, 21, 16, 17, 22, 23, 24, 18, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34};
__device__ const signed char ACCF_ROW[ACC_N][3] = {
    {-1, -1, -1}, {-1, -1, -1}, {-1, -1, -1}, {-1, -1, -1}, {1, 2, -1}, {0, 1, 2}, {0, 1, 2},
    {-1, -1, -1}, {-1, -1, -1}, {-1, -1, -1}, {-1, -1, -1}, {-1, -1, -1}, {-1, -1, -1}, {-1, -1, -1}, {-1, -1, -1},
    {-1, -1, -1}, {-1, -1, -1}, {-1, -1, -1}, {-1, -1, -1}, {-1, -1, -1}, {-1, -1, -1}, {-1, -1, -1},
    {1, 2, -1}, {1, 2, -1}, {1, 2, -1}, {0, 1, 2}, {0, 1, 2}, {0, 1, 2}, {-1, -1, -1}};
__device__ const signed char ACCF_VEC[ACC_N][3] = {
    {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 1, 0}, {2, 3, 4}, {5, 6, 7},
    {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0},
    {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0},
    {8, 9, 0}, {10, 11, 0}, {12, 13, 0}, {14, 15, 16}, {17, 18, 19}, {20, 21, 22}, {0, 0, 0}};

__device__ __forceinline__ double acc_finish(const double *t /*NACC*/, const AngTab &ang, int o) {
    const double *tab = &ang.j[0][0];
    const int base = ACCF_BASE[o];
    double v = (base >= 0) ? t[base] : 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int r = ACCF_ROW[o][k];
        if (r >= 0) {
            const double *P = t + 4 + 3 * r;
            const double *h = tab + 3 * ACCF_VEC[o][k];
            v += P[0] * h[0] + P[1] * h[1] + P[2] * h[2];
        }
    }
    return v;
}

// Warp-cooperative version of lu_solve6 (b2_ndt_math.cuh): lane r < 6 owns row r of the augmented matrix
// [H^T | -g]; pivot search, row exchange and elimination use shuffles, every operation on a matrix entry is
// the one the serial code performs (same pivot rule: first row with the largest |entry|; same f = a*inv,
// a -= f*p arithmetic), so the solution is the serial one.  All lanes return the solution and the
// min|pivot| / max|pivot| estimate (0 when a pivot is zero or NaN).
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

__device__ __noinline__ double warp_lu_solve6(const double *__restrict__ H, const double *__restrict__ g, int lane, double x[6]) {
    const int r = (lane < 6) ? lane : 5;
    double a[7];
#pragma unroll
    for (int c = 0; c < 6; ++c) a[c] = H[c * 6 + r];
    a[6] = -g[r];
    double pmin = DBL_MAX, pmax = 0.0;
    bool singular = false;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double mine = fabs(a[k]);
        double best = shfl_d(mine, k);
        int pl = k;
#pragma unroll
        for (int rr = k + 1; rr < 6; ++rr) {
            const double v = shfl_d(mine, rr);
            if (v > best) { best = v; pl = rr; }
        }
        if (!(best > 0.0)) singular = true;   // zero or NaN pivot
        pmin = fmin(pmin, best);
        pmax = fmax(pmax, best);
        double pr[7];
#pragma unroll
        for (int c = k; c < 7; ++c) {
            pr[c] = shfl_d(a[c], pl);
            const double kr = shfl_d(a[c], k);
            if (lane == pl) a[c] = kr;        // row exchange k <-> pl (no-op when pl == k)
            if (lane == k) a[c] = pr[c];
        }
        const double inv = 1.0 / pr[k];
        if (lane > k && lane < 6) {
            const double f = a[k] * inv;
#pragma unroll
            for (int c = k + 1; c < 7; ++c) a[c] -= f * pr[c];
        }
    }
    // reciprocal of this lane's pivot: the six divisions run side by side (lu_solve6 multiplies by the same reciprocals)
    const double diag = (r == 0) ? a[0] : (r == 1) ? a[1] : (r == 2) ? a[2] : (r == 3) ? a[3] : (r == 4) ? a[4] : a[5];
    const double rinv = 1.0 / diag;
    double xs[6];
#pragma unroll
    for (int rr = 5; rr >= 0; --rr) {
        double sacc = a[6];
#pragma unroll
        for (int c = rr + 1; c < 6; ++c) sacc -= a[c] * xs[c];
        xs[rr] = shfl_d(sacc * rinv, rr);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = xs[i];
    return singular ? 0.0 : pmin / pmax;
}

// a warp completes a pass request: the twelve sin/cos evaluations run on twelve lanes
__device__ __forceinline__ void finish_request_warp(Slot &S, int lane) {
    __syncwarp();
    if (S.ctl.need_trig) {
        if (lane < 12) {
            const int k = lane % 3;
            const double a = S.ctl.x_req[3 + k];
            if (lane < 3) S.trig_f[k] = sin_f32((float)a);
            else if (lane < 6) S.trig_f[3 + k] = cos_f32((float)a);
            else if (lane < 9) S.trig_d[k] = ang_sin(a);
            else S.trig_d[3 + k] = ang_cos(a);
        }
        __syncwarp();
        // pose matrix + angle tables: three lanes, a third of the tables each
        if (lane < 3) ctl_finish_request(S.ctl, &S.trig_f[0], &S.trig_f[3], &S.trig_d[0], &S.trig_d[3], lane);
        __syncwarp();
        if (lane == 0) S.ctl.need_trig = 0;
    }
    __syncwarp();
}

// One warp runs the controller of a slot on the reduced sums S.raw_total: contraction with the angle tables,
// serial halves on lane 0, the 6x6 Newton solve on the whole warp, then the next pass request.  Returns 1
// when another pass was requested (S.ctl.T / ang / hess describe it), 0 when the match is finished.
#ifdef NDT_TIMING
#define CTL_LAP(k) do { if (ctl_trace && lane == 0) ctl_trace[k] = (unsigned long long)clock64(); } while (0)
#else
#define CTL_LAP(k) do { } while (0)
#endif
__device__ __forceinline__ int controller_step(Slot &S, const NdtConst &K, int deriv_only, int lane, unsigned long long *ctl_trace = nullptr) {
    (void)ctl_trace;
    CTL_LAP(0);
    if (lane < ACC_N) S.total[lane] = acc_finish(S.raw_total, S.ctl.ang, lane);
    __syncwarp();
    CTL_LAP(1);
    int code = CTL_DONE;
    if (!deriv_only) {
        if (lane == 0) code = ctl_pre(S.ctl, K, S.total);
        code = __shfl_sync(0xffffffffu, code, 0);
        CTL_LAP(2);
        for (int guard = 0; guard < 8 && code == CTL_NEWTON; ++guard) {
            __syncwarp();
            double delta[6];
            const double rc = warp_lu_solve6(S.ctl.H, S.ctl.g, lane, delta);
            CTL_LAP(3);
            {
                // the normalised direction delta / |delta| (ctl_post_newton): six divisions on six lanes.  Written before
                // lane 0 decides whether the LU result stands; the SVD fallback recomputes it serially.
                const double nrm = ctl_delta_norm(delta);
                const double dl = (lane == 0) ? delta[0] : (lane == 1) ? delta[1] : (lane == 2) ? delta[2] : (lane == 3) ? delta[3] : (lane == 4) ? delta[4] : delta[5];
                if (lane < 6) S.ctl.dir[lane] = dl / nrm;
                __syncwarp();
            }
            if (lane == 0) {
                bool fin = true;
#pragma unroll
                for (int i = 0; i < 6; ++i) fin = fin && (delta[i] == delta[i]) && (fabs(delta[i]) <= DBL_MAX);
                bool have_dir = true;
                if (K.force_svd || !(rc > 1e-9 && fin)) {
                    // (near-)singular Hessian: Eigen's JacobiSVD solve with its rank truncation
                    double neg_g[6];
                    for (int i = 0; i < 6; ++i) neg_g[i] = -S.ctl.g[i];
                    svd_solve6(S.ctl.H, neg_g, delta);
                    have_dir = false;
                }
                code = ctl_post_newton(S.ctl, K, delta, have_dir);
            }
            code = __shfl_sync(0xffffffffu, code, 0);
            CTL_LAP(4);
        }
        if (code == CTL_NEWTON) { code = CTL_DONE; if (lane == 0) S.ctl.state = ST_DONE; }
    }
    const int go = (code == CTL_PASS) ? 1 : 0;
    __syncwarp();
    if (go) finish_request_warp(S, lane);
    CTL_LAP(5);
    return go;
}

__device__ __forceinline__ void write_result(const Slot &S, const MatchArgs &A, size_t match) {
    for (int i = 0; i < 16; ++i) A.poses_out[match * 16 + i] = S.ctl.finalT[i];
    if (A.results) {
        b2ndt_result r;
        r.iterations = S.ctl.nr_iter; r.converged = S.ctl.converged;
        r.score = S.ctl.score; r.trans_probability = S.ctl.trans_probability;
        for (int i = 0; i < 6; ++i) r.p[i] = S.ctl.p[i];
        r.passes = S.ctl.passes; r.mt_trials = S.ctl.mt_trials; r.pairs = S.ctl.pairs;
        A.results[match] = r;
    }
}

// ---------------------------------------------------------------- search warps (producers) -----
// One derivative pass of one search warp over its share of the source points [first, last): lane = source
// point; neighbour lists walked warp-cooperatively; hits appended to the warp's ring.  The end of the pass is
// published through finished[sw].
struct SearchState {
    uint32_t my_tail = 0;      // entries produced so far (warp-uniform)
    uint32_t hd_seen = 0;      // last consumer position read (warp-uniform)
    uint32_t pass_id = 0;      // passes produced so far
};

// Point of lane l in the round that starts at b: b + l * lstride.  lstride = 1 (batch kernel): a round is 32 consecutive
// points (coalesced, neighbouring points share list lines in L1); lstride = number of search warps of the cluster
// (single match): every warp takes a uniform sample of the scan -- the source is in voxel order, and rounds of
// consecutive points differ several-fold in candidates (one 32-point round near the sensor held a whole cluster at the
// pass barrier for 10 us, profiles/r2_single_match_trace.txt).
__device__ __forceinline__ void search_pass(NdtSmem &S, const GridView &G, const float4 *__restrict__ src, uint32_t first,
                                            uint32_t last, uint32_t b0, uint32_t stride, uint32_t lstride, const float *T, int sw,
                                            int lane, SearchState &st) {
    (void)first;
    const uint32_t loff = (uint32_t)lane * lstride;
    const uint32_t lt = (1u << lane) - 1u;
    float4 *ring = S.ring[sw];
    float4 *stage_all = S.stage[sw];
    uint32_t my_tail = st.my_tail, hd_seen = st.hd_seen;
    ++st.pass_id;
    // wait until the ring has room for `need` more entries
    auto ring_room = [&](uint32_t need) {
        if (my_tail - hd_seen > RING - need) {
            uint32_t hd = 0;
            if (NDT_POLL_ALL || lane == 0) {
                hd = ld_vol(&S.head[sw]);
                while (my_tail - hd > RING - need) {
#if NDT_ROOM_SLEEP > 0
                    __nanosleep(NDT_ROOM_SLEEP);
#endif
                    hd = ld_vol(&S.head[sw]);
                }
            }
            hd_seen = __shfl_sync(0xffffffffu, hd, 0);
        }
    };
    // Software pipeline over rounds of 32 points: while round r is searched, the neighbour-list header
    // of round r+1 and the source points of round r+2 are in flight.
    //   prepare(b, pt, buf): transform the points of the round starting at b, park (point, transformed
    //   point) in stage buffer `buf`, and issue the header load -> (off, cnt); cnt = IRREGULAR marks a
    //   lane that needs the dense-window fallback (4-cell window or centre cell outside the grid).
    constexpr uint32_t IRREGULAR = 0xffffffffu;
    auto load_pt = [&](uint32_t b) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        // ld.global.cg (L2 only): the sources may still be streaming in from the host while the kernel runs, and a
        // 32-byte sector can hold points of two matches that arrive in different chunks
        if (b < last && b + loff < last) v = __ldcg(&src[b + loff]);
        return v;
    };
    {
            auto prepare = [&](uint32_t b, const float4 pt, int buf, uint32_t &off, uint32_t &cnt) {
                off = 0u; cnt = 0u;
                float tx = 0.f, ty = 0.f, tz = 0.f;
                if (b < last && b + loff < last && G.ok) {
                    transform_f32(T, pt.x, pt.y, pt.z, tx, ty, tz);
                    if (finite3(tx, ty, tz)) {
                        // every cell that can hold a centroid within the radius (centroids may sit up to
                        // G.margin outside their own cell; the slack also covers the rounding of this arithmetic)
                        const float q[3] = {tx, ty, tz};
                        int lo[3];
                        bool regular = true;
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const float mg = G.margin + 1e-6f * fabsf(q[a]);
                            lo[a] = (int)floorf((q[a] - G.res - mg) * G.inv_leaf) - G.min_b[a];
                            const int hi = (int)floorf((q[a] + G.res + mg) * G.inv_leaf) - G.min_b[a];
                            // regular: the window is exactly the 3 cells around an in-grid centre cell
                            regular = regular && (hi - lo[a] == 2) && (lo[a] + 1 >= 0) && (lo[a] + 1 < G.div_b[a]);
                        }
                        if (regular) {
                            const uint2 hd = __ldg(&G.nbr_head[(size_t)(lo[0] + 1) + (size_t)(lo[1] + 1) * G.mul[1] +
                                                                (size_t)(lo[2] + 1) * G.mul[2]]);
                            off = hd.x; cnt = hd.y;
                        } else {
                            cnt = IRREGULAR;
                        }
                    }
                }
                stage_all[buf * 64 + lane] = pt;
                stage_all[buf * 64 + 32 + lane] = make_float4(tx, ty, tz, 0.f);
            };
            // three rounds in flight: r (searched now), r+1 (header loaded, its list lines being pulled into
            // L1), r+2 (points loaded, header load issued); stage slot = round % 3
            uint32_t off, cnt, off_1 = 0, cnt_1 = 0, off_2 = 0, cnt_2 = 0;
            prepare(b0, load_pt(b0), 0, off, cnt);
            prepare(b0 + stride, load_pt(b0 + stride), 1, off_1, cnt_1);
            float4 pt_3 = load_pt(b0 + 2u * stride);
#if NDT_LPF
            // the FIRST round of the pass has no predecessor that pulls its list lines in: pull them all now, so that the walk
            // pays one L2 round trip for the round instead of one per 64-entry step (a single match gives every search warp
            // exactly one round: 7 steps, i.e. most of its search time)
            if (cnt != IRREGULAR) {
                const float4 *lp = G.nbr_list + off;
                for (uint32_t k = 0; k < cnt; k += 8u) asm volatile("prefetch.global.L1 [%0];" ::"l"(lp + k));
                if (cnt) asm volatile("prefetch.global.L1 [%0];" ::"l"(lp + cnt - 1u));
            }
#endif
            int slot = 0;
            for (uint32_t base = b0; base < last; base += stride, slot = (slot == 2) ? 0 : slot + 1) {
                __syncwarp();
#if NDT_LPF
                if (cnt_1 != IRREGULAR) {
                    const float4 *lp = G.nbr_list + off_1;
                    for (uint32_t k = 0; k < cnt_1; k += 8u) asm volatile("prefetch.global.L1 [%0];" ::"l"(lp + k));
                    if (cnt_1) asm volatile("prefetch.global.L1 [%0];" ::"l"(lp + cnt_1 - 1u));
                }
#endif
                {
                    const int slot2 = (slot == 0) ? 2 : slot - 1;       // (slot + 2) % 3
                    prepare(base + 2u * stride, pt_3, slot2, off_2, cnt_2);
                    pt_3 = load_pt(base + 3u * stride);
                }
                const float4 *stage = &stage_all[slot * 64];
                const bool irregular = (cnt == IRREGULAR);
                if (irregular) cnt = 0u;
                // ---- regular lanes: the warp walks the concatenation of its 32 neighbour lists, 32 entries
                // at a time (coalesced runs of 16-byte entries); the owner of entry e is found by a binary
                // search over the inclusive scan of the list lengths
                uint32_t incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
                const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                const uint32_t excl = incl - cnt;
                auto fetch = [&](uint32_t e, int &o) {
                    o = 0;
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) {
                        const uint32_t v = __shfl_sync(0xffffffffu, incl, (o + sft - 1) & 31);
                        if (v <= e) o += sft;
                    }
                    o &= 31;
                    const uint32_t o_excl = __shfl_sync(0xffffffffu, excl, o);
                    const uint32_t o_off = __shfl_sync(0xffffffffu, off, o);
                    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e < total) c = __ldg(&G.nbr_list[o_off + (e - o_excl)]);
                    return c;
                };
                // two batches (64 entries) per step: their owner searches, loads and tests are independent
                // instruction streams the scheduler can interleave
#pragma unroll 1
                for (uint32_t e0 = 0; e0 < total; e0 += 64u) {
                    int oa, ob;
                    const float4 ca = fetch(e0 + lane, oa);
                    const float4 cb = fetch(e0 + 32u + lane, ob);
                    ring_room(64u);
                    const float4 qa = stage[32 + oa], qb = stage[32 + ob];
                    // flann::L2_Simple<float>: (dx*dx + dy*dy) + dz*dz, accepted when < (float)(r*r)
                    const float dxa = __fsub_rn(qa.x, ca.x), dya = __fsub_rn(qa.y, ca.y), dza = __fsub_rn(qa.z, ca.z);
                    const float dxb = __fsub_rn(qb.x, cb.x), dyb = __fsub_rn(qb.y, cb.y), dzb = __fsub_rn(qb.z, cb.z);
                    const float d2a = __fadd_rn(__fadd_rn(__fmul_rn(dxa, dxa), __fmul_rn(dya, dya)), __fmul_rn(dza, dza));
                    const float d2b = __fadd_rn(__fadd_rn(__fmul_rn(dxb, dxb), __fmul_rn(dyb, dyb)), __fmul_rn(dzb, dzb));
                    const bool hita = (e0 + lane < total) && (d2a < G.r2);
                    const bool hitb = (e0 + 32u + lane < total) && (d2b < G.r2);
                    const uint32_t ba = __ballot_sync(0xffffffffu, hita);
                    const uint32_t bb = __ballot_sync(0xffffffffu, hitb);
                    const uint32_t na = __popc(ba);
                    if (hita) {
                        const float4 sp = stage[oa];
                        ring[(my_tail + __popc(ba & lt)) & (RING - 1u)] = make_float4(sp.x, sp.y, sp.z, ca.w);
                    }
                    if (hitb) {
                        const float4 sp = stage[ob];
                        ring[(my_tail + na + __popc(bb & lt)) & (RING - 1u)] = make_float4(sp.x, sp.y, sp.z, cb.w);
                    }
                    if (ba | bb) {
                        my_tail += na + __popc(bb);
                        __syncwarp();
                        if (lane == 0) publish_tail(&S.tail[sw], my_tail);
                    }
                }
                // ---- irregular lanes (rare: a centre cell outside the grid -- e.g. returns above the map's height range --
                // or a query within the margin of a cell face): the dense window of ONE such query at a time is searched
                // by the whole warp, lane = window cell (<= 4 x 4 x 4), so a query costs one or two parallel load rounds
                // instead of a serial walk over its rows (which held a single match's cluster at the pass barrier)
                uint32_t irr_mask = __ballot_sync(0xffffffffu, irregular);
                if (irr_mask) {
                    int lo0 = 0, lo1 = 0, lo2 = 0, nx = 0, ny = 0, nz = 0;
                    if (irregular) {
                        const float4 tq = stage[32 + lane];
                        const float q[3] = {tq.x, tq.y, tq.z};
                        int lo[3], nn[3];
                        bool empty = false;
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const float mg = G.margin + 1e-6f * fabsf(q[a]);
                            int l = (int)floorf((q[a] - G.res - mg) * G.inv_leaf) - G.min_b[a];
                            int h = (int)floorf((q[a] + G.res + mg) * G.inv_leaf) - G.min_b[a];
                            l = max(l, 0); h = min(h, G.div_b[a] - 1);
                            lo[a] = l; nn[a] = min(h - l, 3) + 1;      // the window never exceeds 4 cells (margin << cell)
                            empty = empty || (h < l);
                        }
                        if (!empty) { lo0 = lo[0]; lo1 = lo[1]; lo2 = lo[2]; nx = nn[0]; ny = nn[1]; nz = nn[2]; }
                    }
                    while (irr_mask) {
                        const int srcl = __ffs(irr_mask) - 1;
                        irr_mask &= irr_mask - 1u;
                        const int wx = __shfl_sync(0xffffffffu, nx, srcl), wy = __shfl_sync(0xffffffffu, ny, srcl), wz = __shfl_sync(0xffffffffu, nz, srcl);
                        const int b0x = __shfl_sync(0xffffffffu, lo0, srcl), b0y = __shfl_sync(0xffffffffu, lo1, srcl), b0z = __shfl_sync(0xffffffffu, lo2, srcl);
                        const int total_c = wx * wy * wz;
                        if (total_c == 0) continue;
                        const float4 pt = stage[srcl];
                        const float4 tq = stage[32 + srcl];
                        const size_t wbase = (size_t)b0x + (size_t)b0y * G.mul[1] + (size_t)b0z * G.mul[2];
                        for (int kb = 0; kb < total_c; kb += 32) {
                            ring_room(32u);
                            const int k = kb + lane;
                            float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (k < total_c) {
                                const int iz = k / (wx * wy), r = k - iz * wx * wy, iy = r / wx, ix = r - iy * wx;
                                c = __ldg(G.cells + wbase + (size_t)ix + (size_t)iy * G.mul[1] + (size_t)iz * G.mul[2]);
                            }
                            const float ddx = __fsub_rn(tq.x, c.x), ddy = __fsub_rn(tq.y, c.y), ddz = __fsub_rn(tq.z, c.z);
                            const float d2f = __fadd_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)), __fmul_rn(ddz, ddz));
                            const int code = __float_as_int(c.w);
                            const bool hit = (code > 0) && (d2f < G.r2);
                            const uint32_t bm = __ballot_sync(0xffffffffu, hit);
                            if (hit) ring[(my_tail + __popc(bm & lt)) & (RING - 1u)] = make_float4(pt.x, pt.y, pt.z, __int_as_float(code - 1));
                            if (bm) {
                                my_tail += __popc(bm);
                                __syncwarp();
                                if (lane == 0) publish_tail(&S.tail[sw], my_tail);
                            }
                        }
                    }
                }
                off = off_1; cnt = cnt_1; off_1 = off_2; cnt_1 = cnt_2;
            }
    }
    // publish the end of this warp's stream for this pass
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        st_vol(&S.tail[sw], my_tail);
        st_vol(&S.pass_end[sw][st.pass_id & (PASS_RING - 1u)], my_tail);
        st_vol(&S.pass_dead[sw][st.pass_id & (PASS_RING - 1u)], 0u);
        __threadfence_block();
        st_vol(&S.finished[sw], st.pass_id);
    }
    st.my_tail = my_tail; st.hd_seen = hd_seen;
}

// batch kernel: tell this warp's consumer that the slot whose turn it is has no more work (an empty pass
// carrying the dead marker keeps producer and consumer sequences aligned)
__device__ __forceinline__ void search_dead_marker(NdtSmem &S, int sw, int lane, SearchState &st) {
    ++st.pass_id;
    __syncwarp();
    if (lane == 0) {
        st_vol(&S.pass_end[sw][st.pass_id & (PASS_RING - 1u)], st.my_tail);
        st_vol(&S.pass_dead[sw][st.pass_id & (PASS_RING - 1u)], 1u);
        __threadfence_block();
        st_vol(&S.finished[sw], st.pass_id);
    }
}

// Sum of every accumulator over the 32 lanes -> part[NACC].  Each sum is the XOR-butterfly tree (lane l adds the value of
// lane l ^ 16, then ^ 8, 4, 2, 1): the order every earlier version used, so the bits of a match's sums do not change.  But
// instead of 35 full butterflies (175 double shuffles, every lane ending up with every total) the lanes SPLIT the values at
// each step -- the half of the warp with the step's lane bit clear keeps the lower half of the values and hands over the
// upper half, and vice versa -- so a step moves half of what the previous one did: 18 + 9 + 5 + 3 + 2 = 37 double shuffles,
// and each total ends in exactly one lane, which stores it.  (a + b is commutative bit for bit, so the keeper's
// own + received is the value both partners of the full butterfly would compute.)
template <int N, int H>
__device__ __forceinline__ void halve_step(const double (&in)[N], double (&out)[H], const bool upper, const int xor_mask) {
    static_assert(H == (N + 1) / 2, "halves");
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const double a = in[i];
        const double b = (i + H < N) ? in[i + H] : 0.0;
        const double keep = upper ? b : a, send = upper ? a : b;
        out[i] = keep + __shfl_xor_sync(0xffffffffu, send, xor_mask);
    }
}
__device__ __forceinline__ void warp_reduce_acc(const double (&acc)[NACC], const int lane, double *__restrict__ part) {
    double r18[18], r9[9], r5[5], r3[3], r2[2];
    halve_step<NACC, 18>(acc, r18, (lane & 16) != 0, 16);
    halve_step<18, 9>(r18, r9, (lane & 8) != 0, 8);
    halve_step<9, 5>(r9, r5, (lane & 4) != 0, 4);
    halve_step<5, 3>(r5, r3, (lane & 2) != 0, 2);
    halve_step<3, 2>(r3, r2, (lane & 1) != 0, 1);
    // which accumulators this lane ended up with (an index beyond a level's size is that level's zero padding)
    const int o5 = (lane & 1) ? 2 : 0, o4 = (lane & 2) ? 3 : 0, o3 = (lane & 4) ? 5 : 0, o2 = (lane & 8) ? 9 : 0, o1 = (lane & 16) ? 18 : 0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int k5 = o5 + j, k4 = o4 + k5, k3 = o3 + k4, k1 = o1 + o2 + k3;
        if (k5 < 3 && k4 < 5 && k3 < 9 && k1 < NACC) part[k1] = j ? r2[1] : r2[0];
    }
}

// ---------------------------------------------------------------- compute warps (consumers) ----
// Drain one pass from this compute warp's rings: lane = (point, voxel) pair, fixed round-robin over the
// producers, 32 pairs at a time; the warp's partial sums (warp_reduce_acc: the butterfly tree, fixed order) go to `part`.
struct DrainState {
    uint32_t cpos[NDT_PPC];    // entries consumed per producer
    uint32_t pass_id;          // passes drained so far
};

// Returns false when the "pass" was the marker of a dead slot (batch kernel).  The pass description (ctl.T, ctl.ang,
// ctl.hess) is read only once the producers have delivered data or finished, i.e. after its request was published.
template <int NCW>
__device__ __forceinline__ bool compute_drain(NdtSmem &S, const GridView &G, const NdtConst &K, int warp, int lane,
                                              DrainState &ds, double *__restrict__ part) {
    constexpr int PPC = NDT_NSW / NCW;             // producers per compute warp (<= NDT_PPC, the size of ds.cpos)
    static_assert(NCW >= NDT_NCW && NDT_NSW % NCW == 0, "compute warps per CTA");
    ++ds.pass_id;
    bool hess = false, have_desc = false, slot_dead = false;
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
    {
            uint32_t done_mask = 0;                         // bit k: producer k exhausted for this pass
            int turn = 0;
            while (done_mask != (1u << PPC) - 1u) {
                // fixed round-robin over this warp's producers (deterministic accumulation order)
                const int k = turn;
                turn = (turn + 1 == PPC) ? 0 : turn + 1;
                if (done_mask & (1u << k)) continue;
                const int sw = warp + k * NCW;
                uint32_t pos = 0;
#pragma unroll
                for (int kk = 0; kk < PPC; ++kk) if (kk == k) pos = ds.cpos[kk];
                // wait for a full chunk of 32 pairs, or for the producer to finish the pass
                uint32_t n = 0;
                if (NDT_POLL_ALL || lane == 0) {
                    while (true) {
                        // tail first, finished second: when the pass is not finished yet, everything up to the
                        // tail read before belongs to it; once finished, the pass ends at pass_end (the producer
                        // may already be appending the next pass)
                        const uint32_t tl = ld_vol(&S.tail[sw]);
                        __threadfence_block();
                        const uint32_t fin = ld_vol(&S.finished[sw]);
                        const bool done = (int32_t)(fin - ds.pass_id) >= 0;
                        __threadfence_block();
                        const uint32_t end = done ? ld_vol(&S.pass_end[sw][ds.pass_id & (PASS_RING - 1u)]) : tl;
                        const uint32_t avail = end - pos;
                        if (avail >= 32u) { n = 32u; break; }
                        if (done) {     // final (possibly empty) chunk; bit 30: dead-slot marker
                            n = avail | 0x80000000u | (ld_vol(&S.pass_dead[sw][ds.pass_id & (PASS_RING - 1u)]) ? 0x40000000u : 0u);
                            break;
                        }
#if NDT_DRAIN_SLEEP > 0
                        __nanosleep(NDT_DRAIN_SLEEP);
#endif
                    }
                }
                n = __shfl_sync(0xffffffffu, n, 0);
                const bool final_chunk = (n & 0x80000000u) != 0;
                if (n & 0x40000000u) slot_dead = true;
                n &= 0x3fffffffu;
                __threadfence_block();
                // The slot whose pass this is comes from shared memory (cur_slot, set by the caller), chunk by chunk: kept in a
                // register across this loop it was spilled, and the reload -- a local-memory load that misses the L1 the list
                // and record streams run through -- stalled every chunk for an L2 round trip (9 % of the compute warps' time).
                const Ctl &ctl = S.slot[ld_vol(&S.cur_slot[warp])].ctl;
                const float *T = ctl.T;
                const AngTab &ang = ctl.ang;
                if (!have_desc) { hess = ctl.hess != 0; have_desc = true; }
                if ((uint32_t)lane < n) {
                    const float4 e = S.ring[sw][(pos + lane) & (RING - 1u)];
                    PairRec r;
                    ndt_pair_load(G.gauss + (size_t)__float_as_uint(e.w) * GAUSS_STRIDE, r);
                    ndt_pair_math(true, e.x, e.y, e.z, r, T, ang, K.d1, K.d2, hess, acc);
                }
                pos += n;
#pragma unroll
                for (int kk = 0; kk < PPC; ++kk) if (kk == k) ds.cpos[kk] = pos;
                __syncwarp();
                if (lane == 0 && n) st_vol(&S.head[sw], pos);
                if (final_chunk) done_mask |= (1u << k);
            }
    }
    warp_reduce_acc(acc, lane, part);
    return !slot_dead;
}

// ================================================================ kernel 1: one match per cluster ======
// One thread-block cluster (1..16 CTAs) per match: lowest latency for a single ScanMatch, also the
// derivatives-only entry point.  Passes are separated by CTA / cluster barriers.
template <int NCW>
__global__ void __launch_bounds__((NCW + NDT_NSW) * 32, (NCW > NDT_NCW) ? 1 : NDT_MIN_CTAS) ndt_match_kernel(GridView G, NdtConst K, MatchArgs A) {
    constexpr int THREADS = (NCW + NDT_NSW) * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NdtSmem &S = *reinterpret_cast<NdtSmem *>(smem_raw);
    Slot &SL = S.slot[0];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned C = cluster.num_blocks();
    const unsigned crank = cluster.block_rank();
    const unsigned match = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_compute = warp < NCW;

    uint32_t first = 0, last = A.n_shared;
    if (A.offsets) { first = A.offsets[match]; last = A.offsets[match + 1]; }
    const uint32_t npts = last - first;

    if (tid < NDT_NSW) { S.tail[tid] = 0u; S.head[tid] = 0u; S.finished[tid] = 0u; }
    if (warp == 0) {
        if (lane == 0) {
            if (A.deriv_only) {
                const double *p = A.poses6 + (size_t)match * 6;
                for (int i = 0; i < 6; ++i) SL.ctl.p[i] = SL.ctl.x_t[i] = p[i];
                ctl_request(SL.ctl, p, 1, ST_INIT);
                SL.ctl.passes = 0; SL.ctl.pairs = 0;
            } else {
                ctl_start(SL.ctl, K, A.guesses + (size_t)match * 16, (double)npts);
            }
            SL.go = 1;
        }
        finish_request_warp(SL, lane);
    }
    __syncthreads();

    const uint32_t stride = C * NDT_NSW * 32u;
#ifdef NDT_TIMING
    // single-match trace (tools/ only): per pass, clock64 at the phase boundaries seen by CTA `A.trace_rank`
#define TRC(slot_) do { if (A.timing && crank == 0 && match == 0 && trc_pass < 6 && (tid == trc_tid)) A.timing[32 + trc_pass * 8 + (slot_)] = (unsigned long long)(clock64() - trc_t0); } while (0)
    const long long trc_t0 = clock64();
    int trc_pass = 0;
    const int trc_tid = is_compute ? 0 : NCW * 32;
#else
#define TRC(slot_) do { } while (0)
#endif
    // The two roles never share code after this point (ptxas sizes each branch for its own register
    // budget); they meet at CTA-wide barriers issued from both branches.
    if (!is_compute) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(NDT_REG_SEARCH));
        const int sw = warp - NCW;
        SearchState st;
        while (true) {
            // rounds are dealt out CTA-first (round = sw * C + crank): the source is in voxel order, so consecutive
            // rounds are spatial neighbours with similar hit counts and every CTA gets an even sample of the scan
            TRC(0);
#ifdef NDT_TIMING
            if (A.timing && match == 0 && trc_pass == 2 && lane == 0 && sw == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); A.timing[352 + crank] = g; }
#endif
            search_pass(S, G, A.src, first, last, first + (sw * C + crank), stride, C * NDT_NSW, SL.ctl.T, sw, lane, st);
            TRC(1);
#ifdef NDT_TIMING
            if (A.timing && match == 0 && trc_pass == 2 && lane == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); A.timing[128 + crank * 16 + warp] = g; }
#endif
#ifdef NDT_TIMING
            ++trc_pass;
#endif
            cta_barrier<THREADS>();                          // (1) all pairs of the pass consumed, partials written
            if (C > 1) cluster.sync();
            cta_barrier<THREADS>();                          // (2) totals ready
            cta_barrier<THREADS>();                          // (3) controller done
            if (!ld_vol(reinterpret_cast<const uint32_t *>(&SL.go))) break;
        }
        if (C > 1) cluster.sync();
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(NDT_REG_COMPUTE));
        int parity = 0;
        if (lane == 0) st_vol(&S.cur_slot[warp], 0u);
        __syncwarp();
        DrainState ds;
#pragma unroll
        for (int k = 0; k < NDT_PPC; ++k) ds.cpos[k] = 0;
        ds.pass_id = 0;
        while (true) {
            (void)compute_drain<NCW>(S, G, K, warp, lane, ds, SL.warp_part[warp]);
            TRC(2);
#ifdef NDT_TIMING
            if (A.timing && match == 0 && trc_pass == 2 && lane == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); A.timing[128 + crank * 16 + warp] = g; }
#endif
            cta_barrier<THREADS>();                              // (1)
            TRC(3);
#ifdef NDT_TIMING
            if (A.timing && match == 0 && trc_pass == 1 && tid == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); A.timing[96 + crank] = g; }
#endif
            // ---------------- deterministic reduction: CTA -> cluster (fixed order) ----------------
            if (tid < NACC) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < NCW; ++w) s += SL.warp_part[w][tid];
                SL.cta_part[parity][tid] = s;
            }
            if (C > 1) {
                cluster.sync();
                TRC(4);
#ifdef NDT_TIMING
                if (A.timing && match == 0 && trc_pass == 1 && tid == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); A.timing[112 + crank] = g; }
#endif
                if (tid < NACC) {
                    // all remote (DSMEM) loads are issued before the first one is consumed; summed in rank order
                    double v[16];
#pragma unroll
                    for (unsigned r = 0; r < 16; ++r) v[r] = (r < C) ? *cluster.map_shared_rank(&SL.cta_part[parity][tid], r) : 0.0;
                    double tot = 0.0;
#pragma unroll
                    for (unsigned r = 0; r < 16; ++r) if (r < C) tot += v[r];
                    SL.raw_total[tid] = tot;
                }
            } else {
                if (tid < NACC) SL.raw_total[tid] = SL.cta_part[parity][tid];
            }
            cta_barrier<THREADS>();                              // (2)
            TRC(5);
            // ---------------- controller: Newton step + More-Thuente state machine (warp 0) ----------------
            if (warp == 0) {
#ifdef NDT_TIMING
                const int go = controller_step(SL, K, A.deriv_only, lane, (A.timing && crank == 0 && match == 0 && trc_pass == 2) ? A.timing + 320 : nullptr);
#else
                const int go = controller_step(SL, K, A.deriv_only, lane);
#endif
                if (lane == 0) SL.go = go;
            }
            cta_barrier<THREADS>();                              // (3)
            TRC(6);
#ifdef NDT_TIMING
            ++trc_pass;
#endif
            if (!SL.go) break;
            parity ^= 1;
        }
        if (C > 1) cluster.sync();    // peers may still be reading this CTA's partials
        if (tid == 0 && crank == 0) {
            if (A.deriv_only) {
                for (int i = 0; i < ACC_N; ++i) A.acc_out[(size_t)match * ACC_N + i] = SL.total[i];
                if (A.poses_out) for (int i = 0; i < 16; ++i) A.poses_out[(size_t)match * 16 + i] = SL.ctl.T[i];
            } else {
                write_result(SL, A, match);
            }
        }
    }
}

// ================================================================ kernel 2: batches, NDT_SLOTS matches per CTA =
// Persistent CTAs (one wave), NDT_SLOTS matches in flight per CTA, work fetched from a global counter.  The
// search warps cycle over the slots pass by pass and never wait at a pass boundary: while the compute warps drain
// the tail of slot A's pass and a compute warp runs A's Newton step, the search warps are already producing the
// passes of the other slots.  Compute warps hand their partial sums to the slot's controller warp through a named
// barrier (bar.arrive for the others, bar.sync for the controller) and move on.  Pass requests reach the search
// warps through one mbarrier per slot: the controller publishes the request and arrives; EACH search warp waits on
// the phase on its own (hardware sleep, no polling of shared memory), so a search warp that is done with its share
// of a pass starts on the next slot without waiting for the slowest search warp (the former named barrier joined
// all eight at every pass end: 17.5 % of the stall samples, profiles/r1_ndt_batch_kernel_ncu_full.txt).  Every warp
// walks the same deterministic sequence (slot 0, 1, .., 0, ..; a slot's next item is its next pass or "dead"), and
// a match's sums are accumulated in the same fixed order as in kernel 1, so results do not depend on scheduling.
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "B2_MBAR_WAIT:\n\t"
#if NDT_MBAR_HINT > 0
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
#endif
        "@p bra B2_MBAR_DONE;\n\t"
        "bra B2_MBAR_WAIT;\n"
        "B2_MBAR_DONE:\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity), "r"((uint32_t)NDT_MBAR_HINT) : "memory");
}

__global__ void __launch_bounds__(NDT_THREADS, NDT_MIN_CTAS) ndt_batch_kernel(GridView G, NdtConst K, MatchArgs A, uint32_t B,
                                                                             uint32_t *__restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NdtSmem &S = *reinterpret_cast<NdtSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_compute = warp < NDT_NCW;
    constexpr uint32_t ALL_SLOTS = (1u << NDT_SLOTS) - 1u;

    // start match m in a slot (controller warp, all lanes): returns 1 and requests its first pass, or 0
    auto start_match = [&](Slot &SL, uint32_t m) -> int {
        if (m >= B) return 0;
        if (A.ready) {
            // host-streamed batch: wait until the copy stream has delivered this match's points.  Every copy this
            // waits for was ENQUEUED BEFORE the kernel was launched (align_host), so the wait does not depend on
            // host progress; it is bounded all the same (a lost copy must not hang the GPU): after ~4 s the CTA
            // raises A.ready[1] and retires, and the host reports the call as failed.
            int ok = 1;
            if (lane == 0) {
                unsigned long long t0 = 0;
                uint32_t spins = 0;
                while (*reinterpret_cast<const volatile uint32_t *>(A.ready) <= m) {
                    __nanosleep(2000);
                    if ((++spins & 1023u) == 0u) {
                        unsigned long long now;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        if (t0 == 0) t0 = now;
                        else if (now - t0 > 4000000000ull) { atomicExch(const_cast<uint32_t *>(A.ready) + 1, 1u); ok = 0; break; }
                    }
                }
                __threadfence();
            }
            ok = __shfl_sync(0xffffffffu, ok, 0);
            if (!ok) return 0;
        }
        if (lane == 0) {
            uint32_t first = 0, last = A.n_shared;
            if (A.offsets) { first = A.offsets[m]; last = A.offsets[m + 1]; }
            SL.first = first; SL.last = last; SL.match = m;
            ctl_start(SL.ctl, K, A.guesses + (size_t)m * 16, (double)(last - first));
        }
        finish_request_warp(SL, lane);
        return 1;
    };
    // the next match from the global counter (it starts behind the statically dealt first round)
    auto fetch_match = [&](Slot &SL) -> int {
        uint32_t m = 0;
        if (lane == 0) m = atomicAdd(work_counter, 1u);
        m = __shfl_sync(0xffffffffu, m, 0);
        return start_match(SL, gridDim.x * NDT_SLOTS + m);
    };

    if (tid < NDT_NSW) { S.tail[tid] = 0u; S.head[tid] = 0u; S.finished[tid] = 0u; }
    if (tid < NDT_SLOTS) mbar_init(&S.req_bar[tid], 1u);
    if (warp < NDT_SLOTS) {
        // first round dealt statically.  A batch that fills every slot: CTA b takes matches b * NDT_SLOTS + j (in
        // batch order, so a host-streamed batch can start on its first chunk); a smaller one: slot j of CTA b takes
        // j * grid + b, so it still spreads over all CTAs
        Slot &SL = S.slot[warp];
        const uint32_t m0 = (B >= gridDim.x * NDT_SLOTS) ? blockIdx.x * NDT_SLOTS + (uint32_t)warp : (uint32_t)warp * gridDim.x + blockIdx.x;
        const int ok = start_match(SL, m0);
        if (lane == 0) { SL.seq = ok ? 1u : 0u; SL.dead = ok ? 0u : 1u; }
    }
    __syncthreads();

    if (!is_compute) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(NDT_REG_SEARCH));
        const int sw = warp - NDT_NCW;
        SearchState st;
        TMB_DECL;
        // every warp walks the same sequence: slot 0, 1, .., 0, .. (one code copy: s is a run-time value); per-slot
        // state lives in bit masks: dead, any request seen, parity of the mbarrier phase to wait for
        uint32_t dead_mask = 0u, seen_mask = 0u, par_mask = 0u;
        int s = 0;
        while (dead_mask != ALL_SLOTS) {
            const uint32_t bit = 1u << s;
            if (!(dead_mask & bit)) {
                TMB_LAP(2);
                // the first request of a slot was published before the initial __syncthreads
                if (seen_mask & bit) { mbar_wait(&S.req_bar[s], (par_mask & bit) ? 1u : 0u); par_mask ^= bit; }
                seen_mask |= bit;
                const int item = ld_vol(&S.slot[s].dead) ? 0 : 1;
                TMB_LAP(0);                                    // [0] waiting for the next pass request
                if (item) {
                    Slot &SL = S.slot[s];
                    const uint32_t first = ld_vol(&SL.first), last = ld_vol(&SL.last);
                    search_pass(S, G, A.src, first, last, first + sw * 32u, NDT_NSW * 32u, 1u, SL.ctl.T, sw, lane, st);
                    TMB_LAP(1);                                // [1] producing
                } else {
                    search_dead_marker(S, sw, lane, st);
                    dead_mask |= bit;
                }
            }
            s = (s + 1 == NDT_SLOTS) ? 0 : s + 1;
        }
        TMB_FLUSH(0, sw == 0);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(NDT_REG_COMPUTE));
        TMB_DECL;
        DrainState ds;
#pragma unroll
        for (int k = 0; k < NDT_PPC; ++k) ds.cpos[k] = 0;
        ds.pass_id = 0;
        uint32_t dead_mask = 0u;
        uint32_t cnt_pack = 0u;                 // passes drained per slot, 4 bits each (only the low bits matter)
        int s = 0;
        while (dead_mask != ALL_SLOTS) {
            const int s_now = s;
            s = (s + 1 == NDT_SLOTS) ? 0 : s + 1;
            {
                const int s = s_now;
                if (dead_mask & (1u << s)) continue;
                TMB_LAP(7);
                cnt_pack = (cnt_pack & ~(0xFu << (4 * s))) | ((((cnt_pack >> (4 * s)) + 1u) & 0xFu) << (4 * s));
                const uint32_t seen_s = (cnt_pack >> (4 * s)) & 0xFu;
                Slot &SL = S.slot[s];
                if (lane == 0) st_vol(&S.cur_slot[warp], (uint32_t)s);
                __syncwarp();
                const bool live = compute_drain<NDT_NCW>(S, G, K, warp, lane, ds, SL.warp_part[warp]);
                if (!live) { dead_mask |= 1u << s; continue; }
                TMB_LAP(1);                                    // [1] draining (incl. waiting for chunks)
                // the controller role of a slot rotates over the compute warps pass by pass (every warp knows the
                // slot's pass count), so the Newton steps do not always delay the same ring consumers
                const int ctl_warp = (int)((uint32_t)s + seen_s) % NDT_NCW;
                if (warp != ctl_warp) {
                    // hand the partial sums to the slot's controller warp and move on
                    __syncwarp();
                    asm volatile("bar.arrive %0, %1;" ::"r"(1 + s), "n"(NDT_NCW * 32) : "memory");
                } else {
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + s), "n"(NDT_NCW * 32) : "memory");
                    TMB_LAP(2);                                // [2] controller warp waiting for the other compute warps
                    // ---- controller of slot s: fixed-order reduction, Newton step, next request / next match
                    for (int i = lane; i < NACC; i += 32) {
                        double sum = 0.0;
#pragma unroll
                        for (int w = 0; w < NDT_NCW; ++w) sum += SL.warp_part[w][i];
                        SL.raw_total[i] = sum;
                    }
                    __syncwarp();
                    int go = controller_step(SL, K, 0, lane);
                    TMB_LAP(3);                                // [3] controller
                    if (!go) {
                        if (lane == 0) write_result(SL, A, SL.match);
                        __syncwarp();
                        go = fetch_match(SL);
                        TMB_LAP(4);                            // [4] result + next match
                    }
                    __syncwarp();
                    if (lane == 0 && !go) st_vol(&SL.dead, 1u);
                    // publish: the request written by the lanes of this warp, then one arrival on the slot's mbarrier
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S.req_bar[s]);
                }
            }
        }
        TMB_FLUSH(8, warp == 0);
        TMB_FLUSH(16, warp == NDT_NCW - 1);
#ifdef NDT_TIMING
        if (warp == 0 && lane == 0 && A.timing) atomicAdd(&A.timing[24], 1ull);
#endif
    }
}

// ------------------------------------------------------------------ fitness score -------------
// pcl::Registration::getFitnessScore: per transformed source point the exact nearest target POINT
// (float squared L2), found by expanding Chebyshev rings over the voxel buckets of the target.
struct FitView {
    const float4 *cells;
    const uint32_t *leaf_start;
    const int32_t *leaf_n;
    const float4 *pts_sorted;
    uint32_t N;
    int32_t min_b[3], div_b[3], mul[3];
    float res, inv_leaf;
    int32_t ok;
};
struct PoseArg { float T[16]; };

constexpr int FIT_THREADS = 128;
constexpr int FIT_WARPS = FIT_THREADS / 32;
constexpr int FIT_RMAX = 8;
constexpr int FIT_U = 4;        // groups of 32 window cells whose loads are in flight together

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// One WARP per source point: the ring walk is warp-uniform (all lanes share the query), the points of every
// bucket are split over the lanes, and the minimum is exact whatever the split (min is order independent).
__global__ void __launch_bounds__(FIT_THREADS, 8) fitness_kernel(FitView F, const float4 *__restrict__ src, uint32_t n, PoseArg P,
                                                              double max_range, double *__restrict__ part_sum,
                                                              unsigned long long *__restrict__ part_cnt) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double sum = 0.0;                    // lane 0 of each warp accumulates its points in index order
    unsigned long long cnt = 0;
    for (uint32_t i = blockIdx.x * FIT_WARPS + warp; i < n; i += gridDim.x * FIT_WARPS) {
        const float4 pt = __ldg(&src[i]);
        float qx, qy, qz;
        transform_f32(P.T, pt.x, pt.y, pt.z, qx, qy, qz);
        if (!F.ok || !finite3(qx, qy, qz)) continue;
        const int cx = (int)floorf(qx * F.inv_leaf) - F.min_b[0];
        const int cy = (int)floorf(qy * F.inv_leaf) - F.min_b[1];
        const int cz = (int)floorf(qz * F.inv_leaf) - F.min_b[2];
        float best = FLT_MAX;
        bool done = false;
        // Chebyshev shells r = 1 (the whole 3x3x3 block, the usual case ends there), 2, 3, ...: the lanes test 32
        // cells of the (2r+1)^3 cube at a time (interior cells were visited by the previous shells), so the cell
        // records and bucket descriptors are fetched in parallel; only occupied buckets are then scanned, their
        // points split over the lanes
        for (int r = 1; r <= FIT_RMAX && !done; ++r) {
            // The cells of shell r, enumerated without the interior the previous shells covered and without the layers
            // above / below the grid (a street map is a few cells high: most of a large cube lies outside it): the
            // (2r-1) interior layers contribute their perimeter ring of 8r cells each, the two face layers dz = -r, +r
            // their whole (2r+1)^2 square.  r = 1 is the full 3x3x3 block.
            const int side = 2 * r + 1, ring = 8 * r, sq = side * side;
            int zlo = -(r - 1), zhi = r - 1;
            if (zlo < -cz) zlo = -cz;
            if (zhi > F.div_b[2] - 1 - cz) zhi = F.div_b[2] - 1 - cz;
            const int nzi = (r == 1) ? 0 : ((zhi >= zlo) ? zhi - zlo + 1 : 0);
            const bool f_lo = (r > 1) && (cz - r >= 0) && (cz - r < F.div_b[2]), f_hi = (r > 1) && (cz + r >= 0) && (cz + r < F.div_b[2]);
            const int n_ring = nzi * ring;
            const int total = (r == 1) ? 27 : n_ring + ((int)f_lo + (int)f_hi) * sq;
            // FIT_U groups of 32 cells per step: their cell records, then their bucket descriptors, are independent loads in
            // flight together (a query far from every target point walks all 8 shells, and the two dependent L2 round
            // trips per group of 32 cells were what the slowest query, i.e. the kernel, took: ~110 us)
            for (int c0 = 0; c0 < total; c0 += 32 * FIT_U) {
                uint32_t bs[FIT_U], bn[FIT_U];
                int32_t code[FIT_U];
#pragma unroll
                for (int u = 0; u < FIT_U; ++u) {
                    const int c = c0 + u * 32 + lane;
                    bs[u] = 0; bn[u] = 0; code[u] = 0;
                    if (c < total) {
                        int dx, dy, dz;
                        if (r == 1) {
                            dx = c % 3 - 1; dy = (c / 3) % 3 - 1; dz = c / 9 - 1;
                        } else if (c < n_ring) {
                            const int layer = c / ring, t = c - layer * ring;
                            dz = zlo + layer;
                            if (t < side) { dy = -r; dx = t - r; }
                            else if (t < 2 * side) { dy = r; dx = t - side - r; }
                            else if (t < 3 * side - 2) { dx = -r; dy = t - 2 * side - r + 1; }
                            else { dx = r; dy = t - (3 * side - 2) - r + 1; }
                        } else {
                            const int o = c - n_ring, f = o / sq, q = o - f * sq;
                            dz = (f == 0 && f_lo) ? -r : r;
                            dx = q % side - r; dy = q / side - r;
                        }
                        const int kx = cx + dx, ky = cy + dy, kz = cz + dz;
                        if (kx >= 0 && ky >= 0 && kz >= 0 && kx < F.div_b[0] && ky < F.div_b[1] && kz < F.div_b[2])
                            code[u] = __float_as_int(__ldg(F.cells + (size_t)kx + (size_t)ky * F.mul[1] + (size_t)kz * F.mul[2]).w);
                    }
                }
#pragma unroll
                for (int u = 0; u < FIT_U; ++u) {
                    if (code[u] != 0) {
                        const int j = (code[u] > 0 ? code[u] : -code[u]) - 1;
                        bs[u] = __ldg(&F.leaf_start[j]);
                        bn[u] = (uint32_t)__ldg(&F.leaf_n[j]);
                    }
                }
#pragma unroll
                for (int u = 0; u < FIT_U; ++u) {
                    uint32_t occ = __ballot_sync(0xffffffffu, bn[u] != 0u);
                    while (occ) {
                        const int src_lane = __ffs(occ) - 1;
                        occ &= occ - 1u;
                        const uint32_t s0 = __shfl_sync(0xffffffffu, bs[u], src_lane), n0 = __shfl_sync(0xffffffffu, bn[u], src_lane);
                        for (uint32_t k = s0 + lane; k < s0 + n0; k += 32u) {
                            const float4 p = __ldg(&F.pts_sorted[k]);
                            const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
                            best = fminf(best, __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
                        }
                    }
                }
            }
            best = warp_min(best);
            // everything not yet visited lies at least r cells away from the query
            const double lim = (double)r * (double)F.res * (1.0 - 1e-4);
            if (best < FLT_MAX && (double)best * (1.0 + 1e-5) <= lim * lim) done = true;
            // the shells already cover the whole grid
            if (cx - r <= 0 && cy - r <= 0 && cz - r <= 0 && cx + r >= F.div_b[0] - 1 && cy + r >= F.div_b[1] - 1 &&
                cz + r >= F.div_b[2] - 1)
                done = true;
        }
        if (!done) {   // rare: isolated query, exact brute force over all target points
            for (uint32_t k = lane; k < F.N; k += 32u) {
                const float4 p = __ldg(&F.pts_sorted[k]);
                const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
                best = fminf(best, __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
            }
            best = warp_min(best);
        }
        if (best < FLT_MAX && (double)best <= max_range) { sum += (double)best; ++cnt; }
    }
    __shared__ double ssum[FIT_WARPS];
    __shared__ unsigned long long scnt[FIT_WARPS];
    if (lane == 0) { ssum[warp] = sum; scnt[warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0; unsigned long long c = 0;
        for (int w = 0; w < FIT_WARPS; ++w) { s += ssum[w]; c += scnt[w]; }
        part_sum[blockIdx.x] = s; part_cnt[blockIdx.x] = c;
    }
}

}  // namespace b2

// ------------------------------------------------------------------ C ABI: b2ndt_* -------------
using namespace b2;

struct b2ndt {
    int device = 0;
    b2ndt_params prm;
    NdtConst K;
    cudaStream_t own = nullptr, st = nullptr;
    VoxPipeline pipe;
    TargetDev tgt;
    DevBuf d_src, d_guess, d_pose, d_res, d_off, d_p6, d_acc, d_fit_sum, d_fit_cnt;
    PinBuf h_stage, h_small, h_res;
    size_t last_n = 0;
    const float4 *last_src = nullptr;     // device source of the last ScanMatch (GetFitnessScore)
    float last_pose[16];
    bool have_last = false;
    int cl_single = 16, cl_batch = 0;      // CTAs per match: single ScanMatch (upper bound) / batches (0 = by batch size)
    bool attrs_set = false, batch_attrs_set = false;
    bool use_batch_kernel = true;          // B2NDT_BATCH_KERNEL=0 falls back to one-match-per-CTA launches (A/B testing)
    int batch_ctas = 0;
    DevBuf d_work;                         // [0] next match to fetch, [1] matches whose sources are resident
    cudaStream_t copy_st = nullptr;        // H2D stream of host-streamed batches
    cudaEvent_t ev = nullptr;
    PinBuf h_ready;
    bool stream_batches = true;            // B2NDT_STREAM=0: copy everything, then launch
    bool small_batch_clusters = true;      // B2NDT_SMALL_BATCH=0: never widen the matches of a small batch to clusters
    bool zero_copy_small = true;           // B2NDT_ZERO_COPY=0: guesses / results of small launches through copy operations (A/B runs)
    bool wide_single = true;               // B2NDT_WIDE_SINGLE=0: launches that fit one CTA per SM keep the 4 + 8 warp shape (A/B runs)
};

static void gauss_constants(double outlier_ratio, float resolution, double *d1, double *d2) {
    // NDTM:315-321
    double c1 = 10.0 * (1.0 - outlier_ratio);
    double c2 = outlier_ratio / pow((double)resolution, 3);
    double d3 = -log(c2);
    *d1 = -log(c1 + c2) - d3;
    *d2 = -2.0 * log((-log(c1 * exp(-0.5) + c2) - d3) / *d1);
}

extern "C" void b2ndt_params_default(b2ndt_params *p) {
    if (!p) return;
    p->res = 1.0f; p->step_size = 0.1; p->trans_eps = 0.01; p->outlier_ratio = 0.55;
    p->max_iter = 30; p->min_pts = 6; p->eig_mult = 0.01; p->pcl17_compat = 1;
}

extern "C" int b2ndt_create(const b2ndt_params *p, int device, b2ndt **out) {
    if (!out || !p) { set_error("b2ndt_create: NULL argument"); return B2_ERR_INVALID; }
    *out = nullptr;
    if (!(p->res > 0.f) || !(p->step_size > 0) || p->max_iter < 0 || p->min_pts < 1) {
        set_error("b2ndt_create: invalid parameters (res %g step %g iter %d min_pts %d)", p->res, p->step_size, p->max_iter, p->min_pts);
        return B2_ERR_INVALID;
    }
    int rc = check_device(device);
    if (rc) return rc;
    B2_CUDA(cudaSetDevice(device));
    b2ndt *h = new b2ndt();
    h->device = device;
    h->prm = *p;
    gauss_constants(p->outlier_ratio, p->res, &h->K.d1, &h->K.d2);
    h->K.step_size = p->step_size; h->K.trans_eps = p->trans_eps; h->K.max_iter = p->max_iter;
    h->K.pcl17_compat = p->pcl17_compat; h->K.force_svd = 0; h->K.res = p->res;
    cudaError_t e = cudaStreamCreateWithFlags(&h->own, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e)); delete h; return B2_ERR_CUDA; }
    h->st = h->own;
    {
        // load the two large match kernels now (CUDA loads kernels lazily, ~0.1-0.25 s each on first launch): the
        // cost belongs to construction, not to the first ScanMatch of a 10 Hz pipeline
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, ndt_match_kernel<NDT_NCW>) != cudaSuccess) cudaGetLastError();
        if (cudaFuncGetAttributes(&fa, ndt_match_kernel<NDT_NCW_WIDE>) != cudaSuccess) cudaGetLastError();
        if (cudaFuncGetAttributes(&fa, ndt_batch_kernel) != cudaSuccess) cudaGetLastError();
    }
    if (const char *e = getenv("B2NDT_BATCH_KERNEL")) h->use_batch_kernel = atoi(e) != 0;
    if (const char *e = getenv("B2NDT_STREAM")) h->stream_batches = atoi(e) != 0;
    if (const char *e = getenv("B2NDT_SMALL_BATCH")) h->small_batch_clusters = atoi(e) != 0;
    if (const char *e = getenv("B2NDT_WIDE_SINGLE")) h->wide_single = atoi(e) != 0;
    if (const char *e = getenv("B2NDT_ZERO_COPY")) h->zero_copy_small = atoi(e) != 0;
    *out = h;
    return 0;
}

extern "C" void b2ndt_destroy(b2ndt *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->st);
    h->pipe.release();
    TargetDev &t = h->tgt;
    t.pts_in.release(); t.pts_sorted.release(); t.leaf_idx.release(); t.leaf_n.release(); t.leaf_start.release();
    t.centroid4.release(); t.gauss.release(); t.sums.release(); t.icov9.release(); t.cells.release(); t.counters.release();
    t.nbr_head.release(); t.nbr_list.release(); t.nbr_tiles.release();
    t.csum4.release(); t.touched.release(); t.pts_all.release();
    h->d_src.release(); h->d_guess.release(); h->d_pose.release(); h->d_res.release(); h->d_off.release();
    h->d_work.release(); h->h_ready.release();
    if (h->copy_st) { cudaStreamSynchronize(h->copy_st); cudaStreamDestroy(h->copy_st); }
    if (h->ev) cudaEventDestroy(h->ev);
    h->d_p6.release(); h->d_acc.release(); h->d_fit_sum.release(); h->d_fit_cnt.release();
    h->h_stage.release(); h->h_small.release(); h->h_res.release();
    if (h->own) cudaStreamDestroy(h->own);
    delete h;
}

extern "C" int b2ndt_set_stream(b2ndt *h, void *stream) {
    if (!h) { set_error("b2ndt_set_stream: NULL handle"); return B2_ERR_INVALID; }
    h->st = stream ? (cudaStream_t)stream : h->own;
    return 0;
}

extern "C" int b2ndt_synchronize(b2ndt *h) {
    if (!h) { set_error("b2ndt_synchronize: NULL handle"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(h->device));
    B2_CUDA(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int b2ndt_set_cluster(b2ndt *h, int single_match_ctas, int batch_ctas) {
    if (!h) { set_error("b2ndt_set_cluster: NULL handle"); return B2_ERR_INVALID; }
    if (single_match_ctas < 1 || single_match_ctas > 16 || batch_ctas < 0 || batch_ctas > 16) {
        set_error("b2ndt_set_cluster: cluster width must be 1..16 (batch: 0 = chosen by batch size)"); return B2_ERR_INVALID;
    }
    h->cl_single = single_match_ctas; h->cl_batch = batch_ctas;
    return 0;
}

static int build_target(b2ndt *h, const float4 *d_pts, size_t n, int nbits_hint) {
    TargetDev &t = h->tgt;
    t.valid = false; t.N = (uint32_t)n; t.V = 0; t.n_tree = 0; t.max_disp = 0.f;
    t.stale_buckets = false; t.unsorted = false;
    if (d_pts != t.pts_all.as<float4>()) { t.src = d_pts; t.n_all = n; t.owns_all = false; t.upd_incremental = t.upd_rebuilt = 0; }   // else: a rebuild out of the update path
    // The dense arrays are cleared entry by entry, not as a whole (117 MB for the headline map): zero what the previous
    // build wrote, while its leaf table and listed-cell table are still intact.
    bool dense_zero = false;
    if (t.fp_clean && t.cells.p && t.nbr_head.p) {
        const uint32_t m = t.fp_V > t.fp_active ? t.fp_V : t.fp_active;
        if (m) {
            grid_clear_kernel<<<(m + 255) / 256, 256, 0, h->st>>>(t.leaf_idx.as<int32_t>(), t.fp_V, t.nbr_tiles.as<uint32_t>(), t.fp_active,
                                                               t.cells.as<float4>(), t.nbr_head.as<uint2>());
            B2_LAUNCH_CHECK();
        }
        dense_zero = true;
    }
    t.fp_clean = false; t.fp_V = 0; t.fp_active = 0;
    h->have_last = false;
    memset(&t.L, 0, sizeof(t.L));
    int rc;
    uint32_t off[2] = {0u, (uint32_t)n};
    if ((rc = h->pipe.plan(off, 1, h->st))) return rc;
    const float res = h->prm.res;
    if ((rc = h->pipe.run(d_pts, res, res, res, nbits_hint, h->st))) return rc;
    if ((rc = h->h_small.reserve(4096))) return rc;
    uint32_t *misc = h->h_small.as<uint32_t>();
    // the one host round trip of the build: the array sizes (occupied voxels, dense-grid cells) are data dependent
    B2_CUDA(cudaMemcpyAsync(misc, h->pipe.scalars(), 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaMemcpyAsync(misc + 8, h->pipe.layouts(), sizeof(VoxLayout), cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaStreamSynchronize(h->st));
    memcpy(&t.L, misc + 8, sizeof(VoxLayout));
    if (!t.L.ok) { t.valid = true; t.fp_clean = dense_zero; return 0; }     // empty cloud or PCL's int32 guard: no cells (PCL warns and clears)
    t.V = misc[1];
    const uint32_t V = t.V;
    if ((rc = t.pts_sorted.reserve((n + 1) * sizeof(float4)))) return rc;
    if ((rc = t.leaf_idx.reserve((V + 1) * 4))) return rc;
    if ((rc = t.leaf_n.reserve((V + 1) * 4))) return rc;
    if ((rc = t.leaf_start.reserve((V + 1) * 4))) return rc;
    if ((rc = t.centroid4.reserve((V + 1) * sizeof(float4)))) return rc;
    if ((rc = t.gauss.reserve((size_t)(V + 1) * GAUSS_STRIDE * 8))) return rc;
    if ((rc = t.sums.reserve((size_t)(V + 1) * 72))) return rc;
    if ((rc = t.csum4.reserve((V + 1) * sizeof(float4)))) return rc;
    if ((rc = t.icov9.reserve((size_t)(V + 1) * 72))) return rc;
    const void *p_cells = t.cells.p, *p_head = t.nbr_head.p;
    if ((rc = t.cells.reserve((size_t)t.L.ncells * 16 + 16))) return rc;
    if ((rc = t.counters.reserve(64 * 4))) return rc;
    if ((rc = t.nbr_head.reserve((size_t)t.L.ncells * 8 + 16))) return rc;
    if ((rc = t.nbr_tiles.reserve(((size_t)t.L.ncells + (size_t)n / LS_SEQ + 64) * 4))) return rc;  // work lists: cells with a list | crowded voxels
    // every searchable leaf enters at most 27 lists: no second round trip to size the lists
    if ((rc = t.nbr_list.reserve(((size_t)(n / (size_t)(h->prm.min_pts > 0 ? h->prm.min_pts : 1) + 1) * 27 + 1) * sizeof(float4)))) return rc;
    // a fresh (or not provably clean) allocation is zeroed once, over its whole capacity: later builds use more of it
    if (!dense_zero || t.nbr_head.p != p_head) B2_CUDA(cudaMemsetAsync(t.nbr_head.p, 0, t.nbr_head.cap, h->st));
    if (!dense_zero || t.cells.p != p_cells) B2_CUDA(cudaMemsetAsync(t.cells.p, 0, t.cells.cap, h->st));
    B2_CUDA(cudaMemsetAsync(t.counters.p, 0, 64 * 4, h->st));
    LayoutArg LA;
    for (int a = 0; a < 3; ++a) { LA.min_b[a] = t.L.min_b[a]; LA.div_b[a] = t.L.div_b[a]; LA.inv[a] = t.L.inv[a]; }
    LA.res = h->prm.res;
    uint32_t *cnt = t.counters.as<uint32_t>();
    uint32_t *active = t.nbr_tiles.as<uint32_t>();
    uint32_t *crowded = active + t.L.ncells;
    if (V) {
        LeafOut O;
        O.pts_sorted = t.pts_sorted.as<float4>(); O.leaf_idx = t.leaf_idx.as<int32_t>(); O.leaf_n = t.leaf_n.as<int32_t>();
        O.leaf_start = t.leaf_start.as<uint32_t>(); O.centroid4 = t.centroid4.as<float4>(); O.sums = t.sums.as<double>();
        O.csum4 = t.csum4.as<float4>();
        LeafUpd U;
        memset(&U, 0, sizeof(U));
        unsigned blocks = ((V / 256 + 147) / 148) * 148u;
        if (blocks < 148) blocks = 148;
        if (blocks > 148 * 16) blocks = 148 * 16;
        leaf_stats_kernel<false><<<blocks, 256, 0, h->st>>>(d_pts, h->pipe.view(), h->pipe.run_start(), V, t.L.n_finite, O, crowded, cnt + 5, U);
        B2_LAUNCH_CHECK();
        launch_chain(leaf_crowded_kernel<false>, 148 * 8, LCROWD_THREADS, 0, h->st, d_pts, h->pipe.view(), h->pipe.run_start(), V, t.L.n_finite, O, crowded, cnt + 5, U);
        B2_LAUNCH_CHECK();
        launch_chain(leaf_finish_kernel, (V + 127) / 128, 128, 0, h->st, V, h->prm.min_pts, h->prm.eig_mult, LA, t.leaf_idx.as<int32_t>(),
                     t.leaf_n.as<int32_t>(), t.centroid4.as<float4>(), t.sums.as<double>(),
                     t.gauss.as<double>(), t.icov9.as<double>(), t.cells.as<float4>(),
                     t.nbr_head.as<uint2>(), cnt, active, (const uint32_t *)nullptr);
        B2_LAUNCH_CHECK();
        // neighbour lists of the listed cells
        launch_chain(nbr_fill_kernel, 148 * 8, NBR_TPB, 0, h->st, t.nbr_head.as<uint2>(), t.cells.as<float4>(), cnt, active, LA, t.nbr_list.as<float4>());
        B2_LAUNCH_CHECK();
    }
    B2_CUDA(cudaMemcpyAsync(misc, t.counters.p, 32, cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaStreamSynchronize(h->st));
    t.n_tree = misc[0];
    memcpy(&t.max_disp, &misc[1], 4);
    if (misc[4] >= NB_MASK) { set_error("SetInputTarget: %u neighbour-list entries exceed the 2^30 limit", misc[4]); return B2_ERR_INVALID; }
    t.fp_V = V; t.fp_active = misc[3]; t.fp_clean = true;
    t.valid = true;     // only a completely built target is usable: a failed allocation above leaves "no target set"
    return 0;
}

// grow a device buffer keeping its first `keep` bytes
static int grow_keep(DevBuf &b, size_t bytes, size_t keep, cudaStream_t st) {
    if (bytes <= b.cap) return 0;
    DevBuf nb;
    int rc = nb.reserve(bytes);
    if (rc) return rc;
    if (keep && b.p) B2_CUDA(cudaMemcpyAsync(nb.p, b.p, keep, cudaMemcpyDeviceToDevice, st));
    B2_CUDA(cudaStreamSynchronize(st));
    b.release();
    b = nb;
    return 0;
}

// Incremental target update = NormalDistributionsTransform::updateVoxelGrid -> VoxelGrid::update of the reference's
// in-tree NDT (NormalDistributionsTransform.cpp:968-972, VoxelGrid.cpp:545-584: bounds, :736-809 updateVoxelContent: per
// new point  tmp_centroid += p, tmp_cov += p p^T, then mean / covariance / inverse of the touched voxel again).
// Here: the new points are keyed under the EXISTING layout and sorted by voxel; each run continues the stored sums of
// its leaf (or opens a leaf) in input order, only the touched leaves are finished again, and the neighbour lists are
// laid out again.  Because every sum continues in the order a full build over (old points ++ new points) would use,
// the updated target is that full build BIT FOR BIT (leaf numbering aside).  A new point outside the grid's index box
// changes PCL's layout for the united cloud: the target is then rebuilt from all points (same result by definition).
static int update_target(b2ndt *h, const float4 *d_new, size_t n) {
    TargetDev &t = h->tgt;
    if (!t.valid) { set_error("b2ndt_update_target: no target set"); return B2_ERR_STATE; }
    if (n == 0) return 0;
    if (t.n_all + n >= B2_MAX_POINTS) { set_error("b2ndt_update_target: target too large"); return B2_ERR_INVALID; }
    int rc;
    // the points in hand-over order: an owned copy from the first update on
    if (!t.owns_all) {
        if (t.n_all && !t.src) { set_error("b2ndt_update_target: the target's points are not available"); return B2_ERR_STATE; }
        // (t.src never points into pts_all here: a rebuild out of this path keeps owns_all set)
        if ((rc = t.pts_all.reserve((t.n_all + n + 1) * sizeof(float4)))) return rc;
        if (t.n_all) B2_CUDA(cudaMemcpyAsync(t.pts_all.p, t.src, t.n_all * sizeof(float4), cudaMemcpyDeviceToDevice, h->st));
        t.owns_all = true;
        t.src = t.pts_all.as<float4>();
    } else {
        if ((rc = grow_keep(t.pts_all, (t.n_all + n + 1) * sizeof(float4), t.n_all * sizeof(float4), h->st))) return rc;
        t.src = t.pts_all.as<float4>();
    }
    float4 *d_add = t.pts_all.as<float4>() + t.n_all;
    B2_CUDA(cudaMemcpyAsync(d_add, d_new, n * sizeof(float4), cudaMemcpyDeviceToDevice, h->st));
    t.n_all += n;
    h->have_last = false;
    auto rebuild = [&]() { ++t.upd_rebuilt; return build_target(h, t.pts_all.as<float4>(), t.n_all, 0); };
    if (!t.L.ok || t.V == 0 || t.L.ncells > (1u << 30)) return rebuild();
    // keys of the new points under the existing layout, sorted by voxel, runs = touched voxels
    uint32_t *cnt = t.counters.as<uint32_t>();
    B2_CUDA(cudaMemsetAsync(t.counters.p, 0, 64 * 4, h->st));
    uint32_t off[2] = {0u, (uint32_t)n};
    if ((rc = h->pipe.plan(off, 1, h->st))) return rc;
    {
        unsigned blocks = (unsigned)((n + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        upd_key_kernel<<<blocks, 256, 0, h->st>>>(d_add, (uint32_t)n, t.L, h->pipe.keys0(), cnt + 7, cnt + 8);
        B2_LAUNCH_CHECK();
    }
    if ((rc = h->pipe.run_prepared(t.L.ncells, t.L.nbits, h->st))) return rc;
    if ((rc = h->h_small.reserve(4096))) return rc;
    uint32_t *misc = h->h_small.as<uint32_t>();
    B2_CUDA(cudaMemcpyAsync(misc, h->pipe.scalars(), 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaMemcpyAsync(misc + 8, cnt, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaStreamSynchronize(h->st));
    const uint32_t R = misc[1], n_valid = misc[8 + 7], oob = misc[8 + 8];
    if (oob) return rebuild();
    ++t.upd_incremental;
    if (R == 0) { t.N = (uint32_t)t.n_all; return 0; }            // only non-finite points
    // from here on the tables are modified in place: a failure leaves "no target set" and dense arrays of unknown content
    t.valid = false; t.fp_clean = false;
    const uint32_t V0 = t.V;
    const size_t Vcap = (size_t)V0 + R + 1;
    if ((rc = grow_keep(t.leaf_idx, Vcap * 4, (size_t)V0 * 4, h->st))) return rc;
    if ((rc = grow_keep(t.leaf_n, Vcap * 4, (size_t)V0 * 4, h->st))) return rc;
    if ((rc = grow_keep(t.leaf_start, Vcap * 4, (size_t)V0 * 4, h->st))) return rc;
    if ((rc = grow_keep(t.centroid4, Vcap * sizeof(float4), (size_t)V0 * sizeof(float4), h->st))) return rc;
    if ((rc = grow_keep(t.csum4, Vcap * sizeof(float4), (size_t)V0 * sizeof(float4), h->st))) return rc;
    if ((rc = grow_keep(t.gauss, Vcap * GAUSS_STRIDE * 8, (size_t)V0 * GAUSS_STRIDE * 8, h->st))) return rc;
    if ((rc = grow_keep(t.sums, Vcap * 72, (size_t)V0 * 72, h->st))) return rc;
    if ((rc = grow_keep(t.icov9, Vcap * 72, (size_t)V0 * 72, h->st))) return rc;
    if ((rc = t.touched.reserve(((size_t)R + 1) * 4))) return rc;
    // the listed-cell table must survive (it names the list heads to zero); the crowded-run list behind it and the
    // neighbour lists themselves are rebuilt
    if ((rc = grow_keep(t.nbr_tiles, ((size_t)t.L.ncells + (size_t)n / LS_SEQ + 64) * 4, (size_t)t.fp_active * 4, h->st))) return rc;
    if ((rc = t.nbr_list.reserve(((size_t)(t.n_all / (size_t)(h->prm.min_pts > 0 ? h->prm.min_pts : 1) + 1) * 27 + 1) * sizeof(float4)))) return rc;
    LayoutArg LA;
    for (int a = 0; a < 3; ++a) { LA.min_b[a] = t.L.min_b[a]; LA.div_b[a] = t.L.div_b[a]; LA.inv[a] = t.L.inv[a]; }
    LA.res = h->prm.res;
    uint32_t *active = t.nbr_tiles.as<uint32_t>();
    uint32_t *crowded = active + t.L.ncells;
    LeafOut O;
    O.pts_sorted = nullptr; O.leaf_idx = t.leaf_idx.as<int32_t>(); O.leaf_n = t.leaf_n.as<int32_t>();
    O.leaf_start = t.leaf_start.as<uint32_t>(); O.centroid4 = t.centroid4.as<float4>(); O.sums = t.sums.as<double>();
    O.csum4 = t.csum4.as<float4>();
    LeafUpd U;
    U.cells = t.cells.as<float4>(); U.leaf_n = t.leaf_n.as<int32_t>(); U.V_old = V0; U.new_ctr = cnt + 6; U.touched = t.touched.as<uint32_t>();
    unsigned blocks = ((R / 256 + 147) / 148) * 148u;
    if (blocks < 148) blocks = 148;
    if (blocks > 148 * 16) blocks = 148 * 16;
    leaf_stats_kernel<true><<<blocks, 256, 0, h->st>>>(d_add, h->pipe.view(), h->pipe.run_start(), R, n_valid, O, crowded, cnt + 5, U);
    B2_LAUNCH_CHECK();
    leaf_crowded_kernel<true><<<148 * 8, LCROWD_THREADS, 0, h->st>>>(d_add, h->pipe.view(), h->pipe.run_start(), R, n_valid, O, crowded, cnt + 5, U);
    B2_LAUNCH_CHECK();
    leaf_finish_kernel<<<(R + 127) / 128, 128, 0, h->st>>>(R, h->prm.min_pts, h->prm.eig_mult, LA, t.leaf_idx.as<int32_t>(),
                                                          t.leaf_n.as<int32_t>(), t.centroid4.as<float4>(), t.sums.as<double>(),
                                                          t.gauss.as<double>(), t.icov9.as<double>(), t.cells.as<float4>(),
                                                          t.nbr_head.as<uint2>(), cnt, active, t.touched.as<uint32_t>());
    B2_LAUNCH_CHECK();
    // neighbour lists again, over every leaf: the list heads of the cells listed so far are zeroed, then recounted
    if (t.fp_active) {
        grid_clear_kernel<<<(t.fp_active + 255) / 256, 256, 0, h->st>>>(nullptr, 0u, active, t.fp_active, nullptr, t.nbr_head.as<uint2>());
        B2_LAUNCH_CHECK();
    }
    B2_CUDA(cudaMemcpyAsync(misc, cnt, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaStreamSynchronize(h->st));
    const uint32_t V1 = V0 + misc[6];
    nbr_count_kernel<<<(V1 + 127) / 128, 128, 0, h->st>>>(V1, h->prm.min_pts, LA, t.leaf_idx.as<int32_t>(), t.leaf_n.as<int32_t>(),
                                                         t.centroid4.as<float4>(), t.nbr_head.as<uint2>(), cnt, active);
    B2_LAUNCH_CHECK();
    nbr_fill_kernel<<<148 * 8, NBR_TPB, 0, h->st>>>(t.nbr_head.as<uint2>(), t.cells.as<float4>(), cnt, active, LA, t.nbr_list.as<float4>());
    B2_LAUNCH_CHECK();
    B2_CUDA(cudaMemcpyAsync(misc, cnt, 32, cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaStreamSynchronize(h->st));
    if (misc[4] >= NB_MASK) { set_error("b2ndt_update_target: %u neighbour-list entries exceed the 2^30 limit", misc[4]); return B2_ERR_INVALID; }
    t.fp_V = V1; t.fp_active = misc[3]; t.fp_clean = true;
    t.valid = true;
    t.n_tree = misc[0];
    memcpy(&t.max_disp, &misc[1], 4);
    if (V1 != V0) t.unsorted = true;
    t.V = V1;
    t.N = (uint32_t)t.n_all;
    t.L.n_finite += n_valid;
    t.stale_buckets = true;
    return 0;
}

// CTAs per single match: as many as give every search warp at most one round of 32 source points (a warp with two
// rounds makes every CTA of the cluster wait at the pass barrier), capped by b2ndt_set_cluster; any size 1..16
static int single_match_cluster(const b2ndt *h, size_t n) {
    const size_t rounds = (n + 31) / 32;
    size_t want = (rounds + NDT_NSW - 1) / NDT_NSW;
    if (want < 1) want = 1;
    return (int)(want < (size_t)h->cl_single ? want : (size_t)h->cl_single);
}

static int check_cloud_args(const char *fn, const void *pts, size_t n, size_t stride, size_t ioff) {
    if (n && !pts) { set_error("%s: NULL cloud", fn); return B2_ERR_INVALID; }
    if (stride < 16 || (stride & 3) || ioff + 4 > stride || (ioff & 3)) { set_error("%s: bad stride %zu / intensity offset %zu", fn, stride, ioff); return B2_ERR_INVALID; }
    if (n >= B2_MAX_POINTS) { set_error("%s: cloud too large", fn); return B2_ERR_INVALID; }
    return 0;
}

extern "C" int b2ndt_set_target(b2ndt *h, const void *pts, size_t n, size_t stride, size_t ioff) {
    if (!h) { set_error("b2ndt_set_target: NULL handle"); return B2_ERR_INVALID; }
    int rc = check_cloud_args("b2ndt_set_target", pts, n, stride, ioff);
    if (rc) return rc;
    B2_CUDA(cudaSetDevice(h->device));
    if ((rc = h->h_stage.reserve(n * 16 + 16))) return rc;
    if ((rc = h->tgt.pts_in.reserve(n * 16 + 16))) return rc;
    int nbits_hint = 0;
    if (n) {
        float mn[3], mx[3];
        pack_cloud_f4_bbox(pts, n, stride, ioff, h->h_stage.as<float>(), mn, mx);
        nbits_hint = key_bits_from_bbox(mn, mx, h->prm.res, h->prm.res, h->prm.res);
        B2_CUDA(cudaMemcpyAsync(h->tgt.pts_in.p, h->h_stage.p, n * 16, cudaMemcpyHostToDevice, h->st));
    }
    return build_target(h, h->tgt.pts_in.as<float4>(), n, nbits_hint);
}

extern "C" int b2ndt_set_target_device(b2ndt *h, const void *d_pts_f4, size_t n) {
    if (!h) { set_error("b2ndt_set_target_device: NULL handle"); return B2_ERR_INVALID; }
    if (n && !d_pts_f4) { set_error("b2ndt_set_target_device: NULL cloud"); return B2_ERR_INVALID; }
    if (n >= B2_MAX_POINTS) { set_error("b2ndt_set_target_device: cloud too large"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(h->device));
    return build_target(h, (const float4 *)d_pts_f4, n, 0);
}

extern "C" int b2ndt_update_target(b2ndt *h, const void *pts, size_t n, size_t stride, size_t ioff) {
    if (!h) { set_error("b2ndt_update_target: NULL handle"); return B2_ERR_INVALID; }
    int rc = check_cloud_args("b2ndt_update_target", pts, n, stride, ioff);
    if (rc) return rc;
    if (n == 0) return h->tgt.valid ? 0 : (set_error("b2ndt_update_target: no target set"), B2_ERR_STATE);
    B2_CUDA(cudaSetDevice(h->device));
    if ((rc = h->h_stage.reserve(n * 16 + 16))) return rc;
    if ((rc = h->d_src.reserve(n * 16 + 16))) return rc;
    pack_cloud_f4(pts, n, stride, ioff, h->h_stage.as<float>());
    B2_CUDA(cudaMemcpyAsync(h->d_src.p, h->h_stage.p, n * 16, cudaMemcpyHostToDevice, h->st));
    return update_target(h, h->d_src.as<float4>(), n);
}

extern "C" int b2ndt_update_target_device(b2ndt *h, const void *d_pts_f4, size_t n) {
    if (!h) { set_error("b2ndt_update_target_device: NULL handle"); return B2_ERR_INVALID; }
    if (n && !d_pts_f4) { set_error("b2ndt_update_target_device: NULL cloud"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(h->device));
    return update_target(h, (const float4 *)d_pts_f4, n);
}

extern "C" int b2ndt_target_info_get(b2ndt *h, b2ndt_target_info *info) {
    if (!h || !info) { set_error("b2ndt_target_info_get: NULL argument"); return B2_ERR_INVALID; }
    if (!h->tgt.valid) { set_error("b2ndt_target_info_get: no target set"); return B2_ERR_STATE; }
    const TargetDev &t = h->tgt;
    info->ok = t.L.ok;
    for (int a = 0; a < 3; ++a) { info->min_b[a] = t.L.min_b[a]; info->div_b[a] = t.L.div_b[a]; }
    info->n_points = t.L.n_finite; info->n_leaves = t.V; info->n_tree = t.n_tree; info->inv_leaf = t.L.inv[0];
    info->updates_incremental = t.upd_incremental; info->updates_rebuilt = t.upd_rebuilt;
    return 0;
}

extern "C" int b2ndt_target_leaves(b2ndt *h, int32_t *idx, int32_t *n_raw, float *centroid4, double *mean3, double *icov9) {
    if (!h) { set_error("b2ndt_target_leaves: NULL handle"); return B2_ERR_INVALID; }
    if (!h->tgt.valid) { set_error("b2ndt_target_leaves: no target set"); return B2_ERR_STATE; }
    const TargetDev &t = h->tgt;
    const size_t V = t.V;
    if (!V) return 0;
    B2_CUDA(cudaSetDevice(h->device));
    B2_CUDA(cudaStreamSynchronize(h->st));
    if (idx) B2_CUDA(cudaMemcpy(idx, t.leaf_idx.p, V * 4, cudaMemcpyDeviceToHost));
    if (n_raw) B2_CUDA(cudaMemcpy(n_raw, t.leaf_n.p, V * 4, cudaMemcpyDeviceToHost));
    if (centroid4) B2_CUDA(cudaMemcpy(centroid4, t.centroid4.p, V * 16, cudaMemcpyDeviceToHost));
    if (icov9) B2_CUDA(cudaMemcpy(icov9, t.icov9.p, V * 72, cudaMemcpyDeviceToHost));
    if (mean3) {
        std::vector<double> g(V * GAUSS_STRIDE);
        B2_CUDA(cudaMemcpy(g.data(), t.gauss.p, V * GAUSS_STRIDE * 8, cudaMemcpyDeviceToHost));
        for (size_t j = 0; j < V; ++j) {
            mean3[3 * j] = g[GAUSS_STRIDE * j]; mean3[3 * j + 1] = g[GAUSS_STRIDE * j + 1]; mean3[3 * j + 2] = g[GAUSS_STRIDE * j + 2];
        }
    }
    if (t.unsorted) {
        // leaves opened by incremental updates sit behind the original ones: report in ascending voxel order all the same
        std::vector<int32_t> key(V);
        B2_CUDA(cudaMemcpy(key.data(), t.leaf_idx.p, V * 4, cudaMemcpyDeviceToHost));
        std::vector<uint32_t> perm(V);
        for (size_t j = 0; j < V; ++j) perm[j] = (uint32_t)j;
        std::sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
        auto apply = [&](auto *arr, size_t width) {
            if (!arr) return;
            std::vector<typename std::remove_pointer<decltype(arr)>::type> tmp(arr, arr + V * width);
            for (size_t j = 0; j < V; ++j)
                for (size_t k = 0; k < width; ++k) arr[j * width + k] = tmp[(size_t)perm[j] * width + k];
        };
        apply(idx, 1); apply(n_raw, 1); apply(centroid4, 4); apply(mean3, 3); apply(icov9, 9);
    }
    return 0;
}

static GridView make_grid_view(const b2ndt *h) {
    GridView G;
    memset(&G, 0, sizeof(G));
    const TargetDev &t = h->tgt;
    G.cells = t.cells.as<float4>();
    G.gauss = t.gauss.as<double>();
    G.nbr_head = t.nbr_head.as<uint2>();
    G.nbr_list = t.nbr_list.as<float4>();
    for (int a = 0; a < 3; ++a) { G.min_b[a] = t.L.min_b[a]; G.div_b[a] = t.L.div_b[a]; G.mul[a] = t.L.mul[a]; }
    G.res = h->prm.res;
    G.r2 = (float)((double)h->prm.res * (double)h->prm.res);
    G.inv_leaf = 1.0f / h->prm.res;
    G.margin = t.max_disp * 1.0001f + 1e-5f * h->prm.res;
    G.ok = (t.L.ok && t.V > 0) ? 1 : 0;
    return G;
}

static int launch_match(b2ndt *h, const MatchArgs &A, size_t B, int C) {
    if (B == 0) return 0;
    if (!h->attrs_set) {
        B2_CUDA(cudaFuncSetAttribute(ndt_match_kernel<NDT_NCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NdtSmem)));
        B2_CUDA(cudaFuncSetAttribute(ndt_match_kernel<NDT_NCW>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        if (NDT_NCW_WIDE != NDT_NCW) {
            B2_CUDA(cudaFuncSetAttribute(ndt_match_kernel<NDT_NCW_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NdtSmem)));
            B2_CUDA(cudaFuncSetAttribute(ndt_match_kernel<NDT_NCW_WIDE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        }
        h->attrs_set = true;
    }
    GridView G = make_grid_view(h);
    if (!h->batch_attrs_set) {
        B2_CUDA(cudaFuncSetAttribute(ndt_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NdtSmem)));
        // one wave of resident CTAs: NDT_MIN_CTAS per SM (__launch_bounds__); CTAs do not depend on one
        // another (work comes from a counter), so a CTA that is not resident at once just starts later
        int sms = 0;
        B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
        h->batch_ctas = NDT_MIN_CTAS * sms;
        h->batch_attrs_set = true;
    }
    if (C == 0) {
        // CTAs per match chosen by batch size.  The persistent kernel keeps 2 matches per CTA in flight, i.e. it needs
        // ~600 matches to fill the GPU and its time is then the latency of two interleaved single-CTA matches; smaller
        // batches (a shard of config 4 or 5 on 8 GPUs: 500 frames / 128 hypotheses) finish sooner when every match is
        // spread over a cluster of 2 / 4 / 8 CTAs, even if the launch then takes up to ~2 waves of CTAs (measured,
        // tools/sweep_small_batches.sh: 500 matches 2.75 -> 2.18 ms with 2 CTAs, 128 matches 1.31 -> 1.10 ms with 4;
        // from ~1000 matches on the persistent kernel wins).
        C = 1;
        if (!A.deriv_only && B >= 1 && h->small_batch_clusters) {
            const size_t two_waves = 2u * (size_t)h->batch_ctas;
            if (B * 8 <= two_waves) C = 8;
            else if (B * 4 <= two_waves) C = 4;
            else if (B <= 800) C = 2;
        }
    }
    if (C == 1 && !A.deriv_only && B >= 2 && h->use_batch_kernel) {
        // batches: persistent CTAs (one wave), NDT_SLOTS matches in flight per CTA, work fetched from a counter
        int rc;
        if ((rc = h->d_work.reserve(64))) return rc;
        if (!A.ready) B2_CUDA(cudaMemsetAsync(h->d_work.p, 0, 8, h->st));     // streamed batches: the caller zeroed it
        // a batch smaller than the number of resident CTAs gets one CTA per match (the first round of matches is dealt
        // CTA-major, so the other slots of those CTAs simply stay empty)
        const unsigned grid = (unsigned)(B < (size_t)h->batch_ctas ? B : (size_t)h->batch_ctas);
        NdtConst Kb = h->K;
        MatchArgs Ab = A;
#ifdef NDT_TIMING
        static unsigned long long *d_tm = nullptr;
        if (!d_tm) cudaMalloc(&d_tm, 32 * 8);
        cudaMemsetAsync(d_tm, 0, 32 * 8, h->st);
        Ab.timing = d_tm;
#endif
        ndt_batch_kernel<<<grid, NDT_THREADS, sizeof(NdtSmem), h->st>>>(G, Kb, Ab, (uint32_t)B, h->d_work.as<uint32_t>());
        b2::count_launch();
        cudaError_t eb = cudaGetLastError();
        if (eb != cudaSuccess) { set_error("ndt_batch_kernel launch failed: %s", cudaGetErrorString(eb)); return B2_ERR_CUDA; }
#ifdef NDT_TIMING
        {
            unsigned long long t[32];
            cudaStreamSynchronize(h->st);
            cudaMemcpy(t, d_tm, sizeof(t), cudaMemcpyDeviceToHost);
            const double n = t[24] ? (double)t[24] : 1.0;
            fprintf(stderr, "[ndt batch timing] ctas %.0f matches %zu | search0: wait %.0f produce %.0f other %.0f | compute0(ctl): wait %.0f drain %.0f barwait %.0f ctl %.0f fetch %.0f other %.0f | compute%d: wait %.0f drain %.0f other %.0f (cycles per CTA)\n",
                    n, B, t[0] / n, t[1] / n, t[2] / n, t[8] / n, t[9] / n, t[10] / n, t[11] / n, t[12] / n, t[15] / n, NDT_NCW - 1,
                    t[16] / n, t[17] / n, t[23] / n);
        }
#endif
        return 0;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(B * (size_t)C));
    // the wide shape (one compute warp per search warp) holds one CTA per SM: taken when every CTA of the launch gets an SM
    // of its own anyway (a single match's cluster, a handful of matches); larger launches keep two CTAs per SM
    const bool wide = (NDT_NCW_WIDE != NDT_NCW) && h->wide_single && (B * (size_t)C <= (size_t)(h->batch_ctas / NDT_MIN_CTAS));
    cfg.blockDim = dim3(wide ? (NDT_NCW_WIDE + NDT_NSW) * 32 : NDT_THREADS);
    cfg.dynamicSmemBytes = sizeof(NdtSmem);
    cfg.stream = h->st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    NdtConst K = h->K;
    MatchArgs Ac = A;
#ifdef NDT_TIMING
    static unsigned long long *d_tr = nullptr;
    if (!d_tr) cudaMalloc(&d_tr, 384 * 8);
    cudaMemsetAsync(d_tr, 0, 384 * 8, h->st);
    Ac.timing = d_tr;
#endif
    cudaError_t e = wide ? cudaLaunchKernelEx(&cfg, ndt_match_kernel<NDT_NCW_WIDE>, G, K, Ac)
                         : cudaLaunchKernelEx(&cfg, ndt_match_kernel<NDT_NCW>, G, K, Ac);
    if (wide && e != cudaSuccess) {
        // the wide shape needs a whole SM per CTA: when such a cluster cannot be placed (SMs held by other work), take the
        // shape that shares an SM
        cudaGetLastError();
        cfg.blockDim = dim3(NDT_THREADS);
        e = cudaLaunchKernelEx(&cfg, ndt_match_kernel<NDT_NCW>, G, K, Ac);
    }
    b2::count_launch();
    if (e != cudaSuccess) { set_error("ndt_match_kernel launch failed: %s", cudaGetErrorString(e)); return B2_ERR_CUDA; }
#ifdef NDT_TIMING
    if (B == 1 && !A.deriv_only) {
        unsigned long long t[384];
        cudaStreamSynchronize(h->st);
        cudaMemcpy(t, d_tr, sizeof(t), cudaMemcpyDeviceToHost);
        {
            // pass 1, every CTA of the cluster on the global timer: arrival at barrier 1 / release from the cluster barrier (ns after the first arrival)
            unsigned long long t0 = ~0ull;
            for (int r = 0; r < C; ++r) if (t[96 + r] && t[96 + r] < t0) t0 = t[96 + r];
            fprintf(stderr, "[ndt cluster] C %d pass 1 b1(ns):", C);
            for (int r = 0; r < C; ++r) fprintf(stderr, " %llu", t[96 + r] - t0);
            fprintf(stderr, " | after cluster barrier:");
            for (int r = 0; r < C; ++r) fprintf(stderr, " %llu", t[112 + r] - t0);
            fprintf(stderr, "\n");
            // pass 2, per CTA: search start (sw 0), then per warp: 4 x compute drained, 8 x search finished (ns after the earliest search start)
            unsigned long long s0 = ~0ull;
            for (int r = 0; r < C; ++r) if (t[352 + r] && t[352 + r] < s0) s0 = t[352 + r];
            fprintf(stderr, "[ndt controller] pass 2 cycles: contraction %llu, ctl_pre %llu, LU solve %llu, post-newton %llu, trig + tables %llu\n",
                    t[321] - t[320], t[322] - t[321], t[323] - t[322], t[324] - t[323], t[325] - t[324]);
            for (int r = 0; r < C; ++r) {
                fprintf(stderr, "[ndt warps] cta %2d start %5llu | drained", r, t[352 + r] - s0);
                const int ncw = wide ? NDT_NCW_WIDE : NDT_NCW;
                for (int w = 0; w < ncw; ++w) fprintf(stderr, " %5llu", t[128 + r * 16 + w] - s0);
                fprintf(stderr, " | search finished");
                for (int w = ncw; w < ncw + NDT_NSW; ++w) fprintf(stderr, " %5llu", t[128 + r * 16 + w] - s0);
                fprintf(stderr, "\n");
            }
        }
        for (int p = 0; p < 3; ++p)
            fprintf(stderr, "[ndt trace] C %d pass %d | search: start %llu finish %llu | compute: drained %llu b1 %llu csync %llu totals(b2) %llu ctl(b3) %llu\n",
                    C, p, t[32 + p * 8 + 0], t[32 + p * 8 + 1], t[32 + p * 8 + 2], t[32 + p * 8 + 3], t[32 + p * 8 + 4], t[32 + p * 8 + 5], t[32 + p * 8 + 6]);
    }
#endif
    return 0;
}

extern "C" int b2ndt_align_batch_device(b2ndt *h, const void *d_src_f4, size_t n_total, const uint32_t *d_offsets, size_t B,
                                        const float *d_guesses, float *d_poses_out, b2ndt_result *d_res) {
    if (!h) { set_error("b2ndt_align_batch_device: NULL handle"); return B2_ERR_INVALID; }
    if (!h->tgt.valid) { set_error("b2ndt_align_batch_device: SetInputTarget has not been called"); return B2_ERR_STATE; }
    if (B == 0) return 0;
    if (!d_guesses || !d_poses_out || (n_total && !d_src_f4)) { set_error("b2ndt_align_batch_device: NULL argument"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(h->device));
    MatchArgs A;
    memset(&A, 0, sizeof(A));
    A.src = (const float4 *)d_src_f4; A.offsets = d_offsets; A.n_shared = (uint32_t)n_total;
    A.guesses = d_guesses; A.poses_out = d_poses_out; A.results = d_res; A.deriv_only = 0;
    return launch_match(h, A, B, h->cl_batch);
}

// is `p` page-locked host memory CUDA can DMA from directly (cudaMallocHost / cudaHostRegister)?
static bool is_pinned_host(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// Host-buffer path of ScanMatch / ScanMatchBatch.  Large batches are pipelined in chunks of matches:
// while the GPU runs the kernel of chunk k the host repacks chunk k+1 and its H2D copy is in flight.
// Packed float4 sources in pinned memory are copied without the staging pass.
static int align_host(b2ndt *h, const void *src, size_t n_total, size_t stride, size_t ioff, const uint32_t *offsets, size_t B,
                      const float *guesses, float *poses_out, b2ndt_result *res, int C) {
    int rc;
    const bool direct = (stride == 16 && ioff == 12 && n_total > 0 && is_pinned_host(src));
    if ((rc = h->h_stage.reserve((direct ? 0 : n_total * 16) + B * 64 + (B + 1) * 4 + 64))) return rc;
    if ((rc = h->d_src.reserve(n_total * 16 + 16))) return rc;
    if ((rc = h->d_guess.reserve(B * 64))) return rc;
    if ((rc = h->d_pose.reserve(B * 64))) return rc;
    if ((rc = h->d_res.reserve(B * sizeof(b2ndt_result)))) return rc;
    if ((rc = h->h_res.reserve(B * (64 + sizeof(b2ndt_result))))) return rc;
    char *stage = h->h_stage.as<char>();
    const size_t src_stage = direct ? 0 : n_total * 16;
    float *gst = (float *)(stage + src_stage);
    memcpy(gst, guesses, B * 64);
    // A handful of matches (a single ScanMatch above all): the kernel reads the guesses from, and writes poses + results
    // to, PINNED HOST memory directly (unified addressing: 64 B read + ~170 B written per match over PCIe) instead of
    // one H2D and two D2H copy operations around it, each of which costs more than the kernel-side access.
    const bool zero_copy = (B <= 16) && h->zero_copy_small;
    if (!zero_copy) B2_CUDA(cudaMemcpyAsync(h->d_guess.p, gst, B * 64, cudaMemcpyHostToDevice, h->st));
    const uint32_t *d_off = nullptr;
    if (offsets) {
        if ((rc = h->d_off.reserve((B + 1) * 4))) return rc;
        uint32_t *ost = (uint32_t *)(stage + src_stage + B * 64);
        memcpy(ost, offsets, (B + 1) * 4);
        B2_CUDA(cudaMemcpyAsync(h->d_off.p, ost, (B + 1) * 4, cudaMemcpyHostToDevice, h->st));
        d_off = h->d_off.as<uint32_t>();
    }
    // Large batches of separate sources: ONE persistent-kernel launch whose sources stream in while it runs.  The
    // copy stream first gets ALL chunk copies (each followed by a bump of the counter of resident matches), then the
    // kernel is launched on the compute stream; a CTA that fetches a match whose points have not arrived yet waits
    // on that counter.  Everything the kernel waits for is already enqueued when it starts, so its progress never
    // depends on the host (CUDA_LAUNCH_BLOCKING, a single hardware queue or a profiler serialising launches cannot
    // dead-lock it), and the wait is bounded on the device.  Strided / unpinned clouds are repacked chunk by chunk
    // before the launch, each chunk's copy overlapping the repack of the next.
    if (offsets && (C == 1 || C == 0) && B >= 256 && h->use_batch_kernel && h->stream_batches) {
        if (!h->copy_st) B2_CUDA(cudaStreamCreateWithFlags(&h->copy_st, cudaStreamNonBlocking));
        if (!h->ev) B2_CUDA(cudaEventCreateWithFlags(&h->ev, cudaEventDisableTiming));
        const size_t nch = 16;
        if ((rc = h->d_work.reserve(64))) return rc;
        if ((rc = h->h_ready.reserve(nch * 4 + 64))) return rc;
        B2_CUDA(cudaMemsetAsync(h->d_work.p, 0, 16, h->st));        // [0] work counter, [1] resident matches, [2] time-out flag
        B2_CUDA(cudaEventRecord(h->ev, h->st));                     // counters zeroed, guesses + offsets queued
        B2_CUDA(cudaStreamWaitEvent(h->copy_st, h->ev, 0));
        uint32_t *hr = h->h_ready.as<uint32_t>();
        const size_t per_c = (B + nch - 1) / nch;
        size_t ci = 0;
        cudaError_t ce = cudaSuccess;
        for (size_t c0 = 0; c0 < B && ce == cudaSuccess; c0 += per_c, ++ci) {
            const size_t c1 = (c0 + per_c < B) ? c0 + per_c : B;
            const size_t p0 = offsets[c0], p1 = offsets[c1];
            if (p1 > p0) {
                const char *from;
                if (direct) from = (const char *)src + p0 * 16;
                else {
                    pack_cloud_f4((const char *)src + p0 * stride, p1 - p0, stride, ioff, (float *)(stage + p0 * 16));
                    from = stage + p0 * 16;
                }
                ce = cudaMemcpyAsync(h->d_src.as<char>() + p0 * 16, from, (p1 - p0) * 16, cudaMemcpyHostToDevice, h->copy_st);
            }
            hr[ci] = (uint32_t)c1;
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(h->d_work.as<uint32_t>() + 1, hr + ci, 4, cudaMemcpyHostToDevice, h->copy_st);
        }
        if (ce != cudaSuccess) {
            cudaStreamSynchronize(h->copy_st);
            cudaStreamSynchronize(h->st);
            set_error("b2ndt_align_batch: streamed H2D copy failed: %s", cudaGetErrorString(ce));
            return B2_ERR_CUDA;
        }
        MatchArgs A;
        memset(&A, 0, sizeof(A));
        A.src = h->d_src.as<float4>(); A.offsets = d_off; A.n_shared = (uint32_t)n_total;
        A.guesses = h->d_guess.as<float>(); A.poses_out = h->d_pose.as<float>(); A.results = h->d_res.as<b2ndt_result>();
        A.ready = h->d_work.as<uint32_t>() + 1;
        if ((rc = launch_match(h, A, B, 1))) { cudaStreamSynchronize(h->copy_st); return rc; }
        char *hres2 = h->h_res.as<char>();
        B2_CUDA(cudaMemcpyAsync(hres2, h->d_pose.p, B * 64, cudaMemcpyDeviceToHost, h->st));
        B2_CUDA(cudaMemcpyAsync(hres2 + B * 64, h->d_res.p, B * sizeof(b2ndt_result), cudaMemcpyDeviceToHost, h->st));
        B2_CUDA(cudaMemcpyAsync(hr + nch, h->d_work.as<uint32_t>() + 2, 4, cudaMemcpyDeviceToHost, h->st));
        B2_CUDA(cudaStreamSynchronize(h->copy_st));
        B2_CUDA(cudaStreamSynchronize(h->st));
        if (hr[nch]) { set_error("b2ndt_align_batch: streamed sources did not arrive on the device (timed out)"); return B2_ERR_CUDA; }
        memcpy(poses_out, hres2, B * 64);
        if (res) memcpy(res, hres2 + B * 64, B * sizeof(b2ndt_result));
        return 0;
    }
    // Chunked launches would overlap the H2D copy with compute, but every extra launch ends in a partial
    // wave that lasts as long as one whole match (~1 ms): measured on B200, 2000 matches take 9.6 ms in one
    // launch and 21 ms in eight.  One launch unless the batch is huge (>= 8 waves per chunk).
    size_t nchunks = 1;
    if (offsets && B >= 16384) nchunks = B / 8192;
    if (const char *e = getenv("B2NDT_CHUNKS")) { int v = atoi(e); if (v >= 1 && offsets) nchunks = (size_t)v; }
    const size_t per = (B + nchunks - 1) / nchunks;
    bool shared_copied = false;
    for (size_t c0 = 0; c0 < B; c0 += per) {
        const size_t c1 = (c0 + per < B) ? c0 + per : B;
        size_t p0 = 0, p1 = n_total;
        if (offsets) { p0 = offsets[c0]; p1 = offsets[c1]; }
        if ((offsets || !shared_copied) && p1 > p0) {
            const char *from;
            if (direct) from = (const char *)src + p0 * 16;
            else {
                pack_cloud_f4((const char *)src + p0 * stride, p1 - p0, stride, ioff, (float *)(stage + p0 * 16));
                from = stage + p0 * 16;
            }
            B2_CUDA(cudaMemcpyAsync(h->d_src.as<char>() + p0 * 16, from, (p1 - p0) * 16, cudaMemcpyHostToDevice, h->st));
            shared_copied = true;
        }
        MatchArgs A;
        memset(&A, 0, sizeof(A));
        A.src = h->d_src.as<float4>(); A.offsets = d_off ? d_off + c0 : nullptr; A.n_shared = (uint32_t)n_total;
        A.guesses = h->d_guess.as<float>() + c0 * 16; A.poses_out = h->d_pose.as<float>() + c0 * 16;
        A.results = h->d_res.as<b2ndt_result>() + c0;
        if (zero_copy) {
            A.guesses = gst + c0 * 16;
            A.poses_out = h->h_res.as<float>() + c0 * 16;
            A.results = reinterpret_cast<b2ndt_result *>(h->h_res.as<char>() + B * 64) + c0;
        }
        if ((rc = launch_match(h, A, c1 - c0, C))) return rc;
    }
    char *hres = h->h_res.as<char>();
    if (!zero_copy) {
        B2_CUDA(cudaMemcpyAsync(hres, h->d_pose.p, B * 64, cudaMemcpyDeviceToHost, h->st));
        B2_CUDA(cudaMemcpyAsync(hres + B * 64, h->d_res.p, B * sizeof(b2ndt_result), cudaMemcpyDeviceToHost, h->st));
    }
    B2_CUDA(cudaStreamSynchronize(h->st));
    memcpy(poses_out, hres, B * 64);
    if (res) memcpy(res, hres + B * 64, B * sizeof(b2ndt_result));
    return 0;
}

extern "C" int b2ndt_align(b2ndt *h, const void *src, size_t n, size_t stride, size_t ioff, const float guess[16], float pose_out[16],
                           b2ndt_result *res) {
    if (!h || !guess || !pose_out) { set_error("b2ndt_align: NULL argument"); return B2_ERR_INVALID; }
    if (!h->tgt.valid) { set_error("b2ndt_align: SetInputTarget has not been called"); return B2_ERR_STATE; }
    int rc = check_cloud_args("b2ndt_align", src, n, stride, ioff);
    if (rc) return rc;
    B2_CUDA(cudaSetDevice(h->device));
    // spread a single match over a cluster only when there is enough work per CTA
    int C = single_match_cluster(h, n);
    rc = align_host(h, src, n, stride, ioff, nullptr, 1, guess, pose_out, res, C < 1 ? 1 : C);
    if (rc) return rc;
    h->last_n = n;
    h->last_src = h->d_src.as<float4>();
    memcpy(h->last_pose, pose_out, 64);
    h->have_last = true;
    return 0;
}

// pcl::transformPointCloud (float, left to right): the align(output) cloud of Registration::align
__global__ void __launch_bounds__(256) transform_cloud_kernel(const float4 *__restrict__ src, uint32_t n, PoseArg P, float4 *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = __ldg(&src[i]);
    float x, y, z;
    transform_f32(P.T, p.x, p.y, p.z, x, y, z);
    out[i] = make_float4(x, y, z, p.w);
}

// b2ndt_align + the result cloud of ScanMatch (registration_interface.hpp:19-22: result_cloud_ptr = the source moved by
// the final pose) filled from the device: the source is already resident, so the transform costs a kernel and a D2H
// instead of a host loop over the points (matters for unfiltered scans).  result_cloud: n records of out_stride bytes
// (PointXYZI: 32 / intensity at 16; data[3] = 1, padding zeroed), may be the source buffer itself.
extern "C" int b2ndt_align_ex(b2ndt *h, const void *src, size_t n, size_t stride, size_t ioff, const float guess[16], float pose_out[16],
                              b2ndt_result *res, void *result_cloud, size_t out_stride, size_t out_ioff) {
    int rc = b2ndt_align(h, src, n, stride, ioff, guess, pose_out, res);
    if (rc || !result_cloud || n == 0) return rc;
    if (out_stride < 16 || (out_stride & 3) || out_ioff + 4 > out_stride || (out_ioff & 3)) { set_error("b2ndt_align_ex: bad output stride / intensity offset"); return B2_ERR_INVALID; }
    if ((rc = h->d_acc.reserve(n * 16 + 16))) return rc;
    if ((rc = h->h_stage.reserve(n * 16 + 16))) return rc;
    PoseArg P;
    memcpy(P.T, pose_out, 64);
    transform_cloud_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->st>>>(h->d_src.as<float4>(), (uint32_t)n, P, h->d_acc.as<float4>());
    B2_LAUNCH_CHECK();
    B2_CUDA(cudaMemcpyAsync(h->h_stage.p, h->d_acc.p, n * 16, cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaStreamSynchronize(h->st));
    const float *r = h->h_stage.as<float>();
    char *o = (char *)result_cloud;
    if (out_stride == 16 && out_ioff == 12) { memcpy(o, r, n * 16); return 0; }
    for (size_t i = 0; i < n; ++i) {
        float *q = (float *)(o + i * out_stride);
        const float x = r[4 * i], y = r[4 * i + 1], z = r[4 * i + 2], w = r[4 * i + 3];
        if (out_stride >= 32) memset(q, 0, out_stride);
        q[0] = x; q[1] = y; q[2] = z;
        if (out_stride >= 32 || out_ioff != 12) q[3] = 1.0f;
        *(float *)(o + i * out_stride + out_ioff) = w;
    }
    return 0;
}

extern "C" int b2ndt_align_batch(b2ndt *h, const void *src, size_t n_total, size_t stride, size_t ioff, const uint32_t *offsets,
                                 size_t B, const float *guesses, float *poses_out, b2ndt_result *res) {
    if (!h || !guesses || !poses_out) { set_error("b2ndt_align_batch: NULL argument"); return B2_ERR_INVALID; }
    if (!h->tgt.valid) { set_error("b2ndt_align_batch: SetInputTarget has not been called"); return B2_ERR_STATE; }
    int rc = check_cloud_args("b2ndt_align_batch", src, n_total, stride, ioff);
    if (rc) return rc;
    if (B == 0) return 0;
    if (offsets) {
        if (offsets[B] != n_total) { set_error("b2ndt_align_batch: offsets[B] != n_total"); return B2_ERR_INVALID; }
        for (size_t b = 0; b < B; ++b) if (offsets[b] > offsets[b + 1]) { set_error("b2ndt_align_batch: offsets not monotone"); return B2_ERR_INVALID; }
    }
    B2_CUDA(cudaSetDevice(h->device));
    h->have_last = false;
    return align_host(h, src, n_total, stride, ioff, offsets, B, guesses, poses_out, res, h->cl_batch);
}

extern "C" int b2ndt_derivatives(b2ndt *h, const void *src, size_t n, size_t stride, size_t ioff, const double p[6], double *score,
                                 double grad[6], double H[36], int64_t *pairs) {
    if (!h || !p) { set_error("b2ndt_derivatives: NULL argument"); return B2_ERR_INVALID; }
    if (!h->tgt.valid) { set_error("b2ndt_derivatives: SetInputTarget has not been called"); return B2_ERR_STATE; }
    int rc = check_cloud_args("b2ndt_derivatives", src, n, stride, ioff);
    if (rc) return rc;
    B2_CUDA(cudaSetDevice(h->device));
    if ((rc = h->h_stage.reserve(n * 16 + 256))) return rc;
    if ((rc = h->d_src.reserve(n * 16 + 16))) return rc;
    if ((rc = h->d_p6.reserve(64))) return rc;
    if ((rc = h->d_acc.reserve(ACC_N * 8))) return rc;
    if ((rc = h->h_res.reserve(ACC_N * 8 + 64))) return rc;
    char *stage = h->h_stage.as<char>();
    if (n) {
        pack_cloud_f4(src, n, stride, ioff, (float *)stage);
        B2_CUDA(cudaMemcpyAsync(h->d_src.p, stage, n * 16, cudaMemcpyHostToDevice, h->st));
    }
    h->have_last = false;      // d_src no longer holds the source of the last ScanMatch (GetFitnessScore needs a new one)
    memcpy(stage + n * 16, p, 48);
    B2_CUDA(cudaMemcpyAsync(h->d_p6.p, stage + n * 16, 48, cudaMemcpyHostToDevice, h->st));
    MatchArgs A;
    memset(&A, 0, sizeof(A));
    A.src = h->d_src.as<float4>(); A.n_shared = (uint32_t)n; A.poses6 = h->d_p6.as<double>();
    A.acc_out = h->d_acc.as<double>(); A.deriv_only = 1;
    int C = single_match_cluster(h, n);
    if ((rc = launch_match(h, A, 1, C < 1 ? 1 : C))) return rc;
    double *acc = h->h_res.as<double>();
    B2_CUDA(cudaMemcpyAsync(acc, h->d_acc.p, ACC_N * 8, cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaStreamSynchronize(h->st));
    if (score) *score = acc[0];
    if (grad) for (int i = 0; i < 6; ++i) grad[i] = acc[1 + i];
    if (H) {
        int k = 7;
        for (int i = 0; i < 6; ++i)
            for (int j = i; j < 6; ++j) { H[j * 6 + i] = acc[k]; H[i * 6 + j] = acc[k]; ++k; }
    }
    if (pairs) *pairs = (int64_t)acc[28];
    return 0;
}

static int fitness_device(b2ndt *h, const float4 *d_src, size_t n, const float pose[16], double max_range, double *out) {
    if (h->tgt.stale_buckets) {
        // incremental updates left the point buckets behind: rebuild from all points (the same target bit for bit)
        const bool had_last = h->have_last;
        int rc = build_target(h, h->tgt.pts_all.as<float4>(), h->tgt.n_all, 0);
        if (rc) return rc;
        h->have_last = had_last;
    }
    const TargetDev &t = h->tgt;
    if (n == 0 || !t.L.ok || t.V == 0) { *out = DBL_MAX; return 0; }
    FitView F;
    memset(&F, 0, sizeof(F));
    F.cells = t.cells.as<float4>(); F.leaf_start = t.leaf_start.as<uint32_t>(); F.leaf_n = t.leaf_n.as<int32_t>();
    F.pts_sorted = t.pts_sorted.as<float4>(); F.N = t.L.n_finite;
    for (int a = 0; a < 3; ++a) { F.min_b[a] = t.L.min_b[a]; F.div_b[a] = t.L.div_b[a]; F.mul[a] = t.L.mul[a]; }
    F.res = h->prm.res; F.inv_leaf = 1.0f / h->prm.res; F.ok = 1;
    PoseArg P;
    memcpy(P.T, pose, 64);
    unsigned blocks = (unsigned)((n + FIT_WARPS - 1) / FIT_WARPS);        // one warp per source point
    if (blocks > 148 * 16) blocks = 148 * 16;
    int rc;
    if ((rc = h->d_fit_sum.reserve(blocks * 8))) return rc;
    if ((rc = h->d_fit_cnt.reserve(blocks * 8))) return rc;
    if ((rc = h->h_res.reserve(blocks * 16 + 64))) return rc;
    fitness_kernel<<<blocks, FIT_THREADS, 0, h->st>>>(F, d_src, (uint32_t)n, P, max_range, h->d_fit_sum.as<double>(),
                                                     h->d_fit_cnt.as<unsigned long long>());
    B2_LAUNCH_CHECK();
    double *hs = h->h_res.as<double>();
    unsigned long long *hc = (unsigned long long *)(hs + blocks);
    B2_CUDA(cudaMemcpyAsync(hs, h->d_fit_sum.p, blocks * 8, cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaMemcpyAsync(hc, h->d_fit_cnt.p, blocks * 8, cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaStreamSynchronize(h->st));
    double s = 0.0;
    unsigned long long c = 0;
    for (unsigned b = 0; b < blocks; ++b) { s += hs[b]; c += hc[b]; }
    *out = c ? s / (double)c : DBL_MAX;
    return 0;
}

extern "C" int b2ndt_fitness(b2ndt *h, double max_range, double *out) {
    if (!h || !out) { set_error("b2ndt_fitness: NULL argument"); return B2_ERR_INVALID; }
    if (!h->tgt.valid || !h->have_last) { set_error("b2ndt_fitness: no completed ScanMatch on this handle"); return B2_ERR_STATE; }
    B2_CUDA(cudaSetDevice(h->device));
    return fitness_device(h, h->last_src, h->last_n, h->last_pose, max_range, out);
}

extern "C" int b2ndt_fitness_ex(b2ndt *h, const void *src, size_t n, size_t stride, size_t ioff, const float pose[16],
                                double max_range, double *out) {
    if (!h || !out || !pose) { set_error("b2ndt_fitness_ex: NULL argument"); return B2_ERR_INVALID; }
    if (!h->tgt.valid) { set_error("b2ndt_fitness_ex: SetInputTarget has not been called"); return B2_ERR_STATE; }
    int rc = check_cloud_args("b2ndt_fitness_ex", src, n, stride, ioff);
    if (rc) return rc;
    B2_CUDA(cudaSetDevice(h->device));
    if ((rc = h->h_stage.reserve(n * 16 + 64))) return rc;
    if ((rc = h->d_src.reserve(n * 16 + 16))) return rc;
    if (n) {
        pack_cloud_f4(src, n, stride, ioff, h->h_stage.as<float>());
        B2_CUDA(cudaMemcpyAsync(h->d_src.p, h->h_stage.p, n * 16, cudaMemcpyHostToDevice, h->st));
    }
    h->have_last = false;
    return fitness_device(h, h->d_src.as<float4>(), n, pose, max_range, out);
}

// ------------------------------------------------------------------ device-resident clouds -----
extern "C" int b2cloud_append_transformed(b2cloud *dst, b2cloud *src, const float T[16]);

// SetInputTarget from a cloud that already lives in HBM (local map assembled / cropped on the device).
extern "C" int b2ndt_set_target_cloud(b2ndt *h, b2cloud *target) {
    if (!h || !target) { set_error("b2ndt_set_target_cloud: NULL argument"); return B2_ERR_INVALID; }
    if (target->device != h->device) { set_error("b2ndt_set_target_cloud: handle and cloud live on different devices"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(h->device));
    return build_target(h, target->d(), target->n, 0);
}

extern "C" int b2ndt_update_target_cloud(b2ndt *h, b2cloud *add) {
    if (!h || !add) { set_error("b2ndt_update_target_cloud: NULL argument"); return B2_ERR_INVALID; }
    if (add->device != h->device) { set_error("b2ndt_update_target_cloud: handle and cloud live on different devices"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(h->device));
    return update_target(h, add->d(), add->n);
}

// ScanMatch with a device-resident source; result_cloud (may be NULL) receives the source transformed by the
// final pose (registration_interface.hpp:19-22).  The source cloud must stay unchanged until the next
// ScanMatch if GetFitnessScore is going to be called (PCL keeps the source pointer the same way).
extern "C" int b2ndt_align_cloud(b2ndt *h, b2cloud *src, const float guess[16], float pose_out[16], b2ndt_result *res,
                                 b2cloud *result_cloud) {
    if (!h || !src || !guess || !pose_out) { set_error("b2ndt_align_cloud: NULL argument"); return B2_ERR_INVALID; }
    if (!h->tgt.valid) { set_error("b2ndt_align_cloud: SetInputTarget has not been called"); return B2_ERR_STATE; }
    if (src->device != h->device) { set_error("b2ndt_align_cloud: handle and cloud live on different devices"); return B2_ERR_INVALID; }
    if (result_cloud == src) { set_error("b2ndt_align_cloud: result_cloud == source"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(h->device));
    int rc;
    const size_t n = src->n;
    if ((rc = h->h_stage.reserve(256))) return rc;
    if ((rc = h->d_guess.reserve(64))) return rc;
    if ((rc = h->d_pose.reserve(64))) return rc;
    if ((rc = h->d_res.reserve(sizeof(b2ndt_result)))) return rc;
    if ((rc = h->h_res.reserve(64 + sizeof(b2ndt_result)))) return rc;
    memcpy(h->h_stage.p, guess, 64);
    B2_CUDA(cudaMemcpyAsync(h->d_guess.p, h->h_stage.p, 64, cudaMemcpyHostToDevice, h->st));
    MatchArgs A;
    memset(&A, 0, sizeof(A));
    A.src = src->d(); A.n_shared = (uint32_t)n; A.guesses = h->d_guess.as<float>(); A.poses_out = h->d_pose.as<float>();
    A.results = h->d_res.as<b2ndt_result>();
    int C = single_match_cluster(h, n);
    if ((rc = launch_match(h, A, 1, C < 1 ? 1 : C))) return rc;
    char *hres = h->h_res.as<char>();
    B2_CUDA(cudaMemcpyAsync(hres, h->d_pose.p, 64, cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaMemcpyAsync(hres + 64, h->d_res.p, sizeof(b2ndt_result), cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaStreamSynchronize(h->st));
    memcpy(pose_out, hres, 64);
    if (res) memcpy(res, hres + 64, sizeof(b2ndt_result));
    h->last_n = n;
    h->last_src = src->d();
    memcpy(h->last_pose, pose_out, 64);
    h->have_last = true;
    if (result_cloud) {
        result_cloud->n = 0;
        if ((rc = b2cloud_append_transformed(result_cloud, src, pose_out))) return rc;
    }
    return 0;
}
