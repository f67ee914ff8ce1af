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
//   gauss       double[10][V]  80-byte records {mean[3], icov xx,xy,xz,yy,yz,zz, pad}
//   cells       float4[ncells] dense grid record {centroid x,y,z, code}: code = +(leaf+1) searchable (n >= min_pts),
//                              -(leaf+1) sparse, 0 empty
// The path is a gather + reduction (no dense contraction): no tensor cores by design.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "b2_common.cuh"
#include "b2_ndt_math.cuh"
#include "b2_voxel.cuh"

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
    bool valid = false;
};

// ------------------------------------------------------------------ target build kernels -----
// One warp per occupied voxel: lanes gather 32 member points (stable = input order), then every lane
// folds them in sequentially via shuffles, reproducing PCL's per-leaf accumulation order exactly:
// mean_ += p (double), cov_ += p p^T (double, from Identity), centroid += (x,y,z,i) (float).
__global__ void __launch_bounds__(256) leaf_stats_kernel(const float4 *__restrict__ pts, const uint32_t *__restrict__ keys,
                                                         const uint32_t *__restrict__ vals,
                                                         const uint32_t *__restrict__ run_start, uint32_t V, uint32_t n_finite,
                                                         float4 *__restrict__ pts_sorted, int32_t *__restrict__ leaf_idx,
                                                         int32_t *__restrict__ leaf_n, uint32_t *__restrict__ leaf_start,
                                                         float4 *__restrict__ centroid4, double *__restrict__ sums) {
    const int l = threadIdx.x & 31;
    for (uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < V; j += (gridDim.x * blockDim.x) >> 5) {
        const uint32_t s = run_start[j];
        const uint32_t e = (j + 1 < V) ? run_start[j + 1] : n_finite;
        float cx = 0.f, cy = 0.f, cz = 0.f, ci = 0.f;
        double sx = 0, sy = 0, sz = 0, cxx = 1, cxy = 0, cxz = 0, cyy = 1, cyz = 0, czz = 1;
        for (uint32_t c = s; c < e; c += 32) {
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c + l < e) { p = __ldg(&pts[vals[c + l]]); pts_sorted[c + l] = p; }
            const int m = (e - c < 32u) ? (int)(e - c) : 32;
#pragma unroll 4
            for (int k = 0; k < m; ++k) {
                float x = __shfl_sync(0xffffffffu, p.x, k), y = __shfl_sync(0xffffffffu, p.y, k);
                float z = __shfl_sync(0xffffffffu, p.z, k), w = __shfl_sync(0xffffffffu, p.w, k);
                cx = __fadd_rn(cx, x); cy = __fadd_rn(cy, y); cz = __fadd_rn(cz, z); ci = __fadd_rn(ci, w);
                double dx = (double)x, dy = (double)y, dz = (double)z;
                sx = __dadd_rn(sx, dx); sy = __dadd_rn(sy, dy); sz = __dadd_rn(sz, dz);
                cxx = __dadd_rn(cxx, __dmul_rn(dx, dx)); cxy = __dadd_rn(cxy, __dmul_rn(dx, dy));
                cxz = __dadd_rn(cxz, __dmul_rn(dx, dz)); cyy = __dadd_rn(cyy, __dmul_rn(dy, dy));
                cyz = __dadd_rn(cyz, __dmul_rn(dy, dz)); czz = __dadd_rn(czz, __dmul_rn(dz, dz));
            }
        }
        if (l == 0) {
            const uint32_t n = e - s;
            const float fn = (float)n;
            leaf_idx[j] = (int32_t)keys[s];
            leaf_n[j] = (int32_t)n;
            leaf_start[j] = s;
            centroid4[j] = make_float4(__fdiv_rn(cx, fn), __fdiv_rn(cy, fn), __fdiv_rn(cz, fn), __fdiv_rn(ci, fn));
            double *o = sums + (size_t)j * 9;
            o[0] = sx; o[1] = sy; o[2] = sz; o[3] = cxx; o[4] = cxy; o[5] = cxz; o[6] = cyy; o[7] = cyz; o[8] = czz;
        }
    }
}

// One thread per voxel: mean, single-pass covariance, eigen inflation, inverse (leaf_finish), the
// 80-byte gather record, and the dense-grid entry.
struct LayoutArg { int32_t min_b[3], div_b[3]; float inv[3]; };

__global__ void __launch_bounds__(128) leaf_finish_kernel(uint32_t V, int min_pts, double eig_mult, LayoutArg LA,
                                                          const int32_t *__restrict__ leaf_idx, const int32_t *__restrict__ leaf_n,
                                                          const float4 *__restrict__ centroid4,
                                                          const double *__restrict__ sums, double *__restrict__ gauss,
                                                          double *__restrict__ icov9, float4 *__restrict__ cells,
                                                          uint32_t *__restrict__ counters) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= V) return;
    const double *s = sums + (size_t)j * 9;
    double sum[3] = {s[0], s[1], s[2]};
    double acc[9] = {s[3], s[4], s[5], s[4], s[6], s[7], s[5], s[7], s[8]};
    double mean[3], cov[9], icov[9], ev[3];
    const int n = leaf_n[j];
    // a leaf that fails PCL's eigenvalue checks (nr_points = -1) stays in the centroid kd-tree with the
    // icov_ it has at that moment (zero, or the non-finite inverse), exactly what leaf_finish leaves here
    (void)leaf_finish(sum, acc, n, min_pts, eig_mult, mean, cov, icov, ev);
    double *g = gauss + (size_t)j * 10;
    g[0] = mean[0]; g[1] = mean[1]; g[2] = mean[2];
    g[3] = icov[0]; g[4] = icov[1]; g[5] = icov[2]; g[6] = icov[4]; g[7] = icov[5]; g[8] = icov[8];
    g[9] = 0.0;
    double *ic = icov9 + (size_t)j * 9;
    for (int a = 0; a < 9; ++a) ic[a] = icov[a];
    const bool tree = n >= min_pts;
    {
        // dense-grid record: the float centroid travels with the voxel code so the search needs one load
        const float4 c = centroid4[j];
        cells[leaf_idx[j]] = make_float4(c.x, c.y, c.z, __int_as_float(tree ? (int32_t)(j + 1) : -(int32_t)(j + 1)));
    }
    if (tree) {
        atomicAdd(&counters[0], 1u);
        // how far the float centroid (PCL's kd-tree point) lies outside its own cell: bounds the search
        // window margin of the match kernel.  Cell k of an axis spans [k/inv, (k+1)/inv).
        const int idx = leaf_idx[j];
        const int iz = idx / (LA.div_b[0] * LA.div_b[1]);
        const int iy = (idx - iz * LA.div_b[0] * LA.div_b[1]) / LA.div_b[0];
        const int ix = idx - iz * LA.div_b[0] * LA.div_b[1] - iy * LA.div_b[0];
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

// ------------------------------------------------------------------ the NDT match kernel ------
// One thread-block cluster (1..16 CTAs) per match; the whole Newton / More-Thuente loop runs inside the
// kernel.  Every CTA is WARP-SPECIALISED (producer / consumer, registers re-balanced with setmaxnreg):
//   search warps  (NDT_NSW, 56 registers): lane = source point.  Float transform, then one z-plane of the
//     3x3x3 window of the dense cell grid per step: nine independent 16-byte loads bring the voxel code
//     AND its float centroid (no second dependent load), float L2 tests, hits appended to the warp's ring
//     buffer in shared memory at positions given by ballots (deterministic order, no atomics);
//   compute warps (NDT_NCW, 128 registers): lane = (point, voxel) pair.  Each compute warp drains the
//     rings of its search warps in a fixed round-robin, 32 pairs at a time, so every lane is busy:
//     80-byte record gather, exp, score / gradient / Hessian terms of updateDerivatives (NDTM:485-520)
//     accumulated in 29 FP64 registers.
// The latency-bound gather and the FP64-bound arithmetic thus run concurrently on different warps, and
// the cheap search warps raise the number of resident warps per SM.
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
constexpr int NDT_WARPS = NDT_NCW + NDT_NSW;
constexpr int NDT_THREADS = NDT_WARPS * 32;
constexpr int NDT_PPC = NDT_NSW / NDT_NCW;     // producers (search warps) per compute warp
constexpr uint32_t RING = 1024;                // ring entries per search warp (power of two, > 512 + 32)
static_assert(NDT_NCW % 4 == 0 && NDT_NSW % 4 == 0 && NDT_NSW % NDT_NCW == 0, "warp-group multiples");

struct GridView {
    const float4 *cells;         // dense grid record {cx, cy, cz, int code}: code > 0 searchable leaf+1, < 0 sparse, 0 empty
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
};

struct NdtSmem {
    Ctl ctl;
    double warp_part[NDT_NCW][ACC_N];
    double cta_part[2][ACC_N];     // double-buffered per-CTA partial, read by cluster peers over DSMEM
    double total[ACC_N];
    double trig_d[6];              // snapped double sin x3, cos x3 of the requested pose
    float  trig_f[6];              // float sin x3, cos x3
    int go;
    int pad_;
    // producer -> consumer rings (monotonic counters, never reset)
    uint32_t tail[NDT_NSW];        // entries produced by search warp s
    uint32_t head[NDT_NSW];        // entries consumed from search warp s
    uint32_t finished[NDT_NSW];    // last pass id search warp s has completed
    uint2 ring[NDT_NSW][RING];     // (source point index, leaf index)
};

__device__ __forceinline__ void cta_barrier() { asm volatile("bar.sync 0, %0;" ::"n"(NDT_THREADS) : "memory"); }
__device__ __forceinline__ uint32_t ld_vol(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void st_vol(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }

__device__ __forceinline__ double dot3v(const double *h, double x, double y, double z) { return x * h[0] + y * h[1] + z * h[2]; }
__device__ __forceinline__ double dot2v(const double *h, double x, double y) { return x * h[0] + y * h[1]; }

// Hessian accumulators: upper triangle, row-major packed, k(i,j) for i<=j:
//   row0: 0..5   row1: 6..10   row2: 11..14   row3: 15,16,17   row4: 18,19   row5: 20      (+7 in acc[])
//
// One (point, voxel) pair: computePointDerivatives (NDTM:448-482) + updateDerivatives (NDTM:485-520).
// J = [I | c3 c4 c5] with c3 = (0,J0,J1), c4 = (J2,J3,J4), c5 = (J5,J6,J7); structural zeros of the angle
// tables (j_ang_f/g/h and h_ang_c/e/f have no z component) are exploited.
__device__ __forceinline__ void ndt_pair(const float px, const float py, const float pz, const float *__restrict__ T,
                                         const AngTab &ang, const double *__restrict__ g, double d1, double d2, bool hess,
                                         double *acc) {
    float tx, ty, tz;
    transform_f32(T, px, py, pz, tx, ty, tz);
    const double2 *g2 = reinterpret_cast<const double2 *>(g);      // 80-byte record, five 16-byte loads
    const double2 a0 = __ldg(g2 + 0), a1 = __ldg(g2 + 1), a2 = __ldg(g2 + 2), a3 = __ldg(g2 + 3), a4 = __ldg(g2 + 4);
    acc[28] += 1.0;
    const double xq = (double)tx - a0.x, yq = (double)ty - a0.y, zq = (double)tz - a1.x;
    const double ixx = a1.y, ixy = a2.x, ixz = a2.y, iyy = a3.x, iyz = a3.y, izz = a4.x;
    const double q0 = ixx * xq + ixy * yq + ixz * zq;
    const double q1 = ixy * xq + iyy * yq + iyz * zq;
    const double q2 = ixz * xq + iyz * yq + izz * zq;
    const double m = xq * q0 + yq * q1 + zq * q2;
    double e = exp(-d2 * m / 2);
    const double sinc = -d1 * e;
    e = d2 * e;
    if (e > 1 || e < 0 || e != e) return;      // NDTM:499-501
    const double w = e * d1;
    acc[0] += sinc;
    const double x = (double)px, y = (double)py, z = (double)pz;
    double J[8];
    J[0] = dot3v(ang.j[0], x, y, z); J[1] = dot3v(ang.j[1], x, y, z);
    J[2] = dot3v(ang.j[2], x, y, z); J[3] = dot3v(ang.j[3], x, y, z); J[4] = dot3v(ang.j[4], x, y, z);
    J[5] = dot2v(ang.j[5], x, y);    J[6] = dot2v(ang.j[6], x, y);    J[7] = dot2v(ang.j[7], x, y);
    // a = J^T q ; gradient += w a
    double a[6];
    a[0] = q0; a[1] = q1; a[2] = q2;
    a[3] = q1 * J[0] + q2 * J[1];
    a[4] = q0 * J[2] + q1 * J[3] + q2 * J[4];
    a[5] = q0 * J[5] + q1 * J[6] + q2 * J[7];
#pragma unroll
    for (int i = 0; i < 6; ++i) acc[1 + i] += w * a[i];
    if (!hess) return;
    double *Hh = &acc[7];
    {   // -d2 w (J^T q)(J^T q)^T
        const double wd = -d2 * w;
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const double t = wd * a[i];
#pragma unroll
            for (int j = i; j < 6; ++j) Hh[k++] += t * a[j];
        }
    }
    {   // w q . H_E(i,j): a=(0,x.a2,x.a3) b=(0,x.b2,x.b3) c=(0,x.c2,x.c3) d=(x.d1,x.d2,x.d3) e=(...) f=(...)
        const double wq0 = w * q0, wq1 = w * q1, wq2 = w * q2;
        Hh[15] += wq1 * dot3v(ang.h[0], x, y, z) + wq2 * dot3v(ang.h[1], x, y, z);
        Hh[16] += wq1 * dot3v(ang.h[2], x, y, z) + wq2 * dot3v(ang.h[3], x, y, z);
        Hh[17] += wq1 * dot2v(ang.h[4], x, y) + wq2 * dot2v(ang.h[5], x, y);
        Hh[18] += wq0 * dot3v(ang.h[6], x, y, z) + wq1 * dot3v(ang.h[7], x, y, z) + wq2 * dot3v(ang.h[8], x, y, z);
        Hh[19] += wq0 * dot2v(ang.h[9], x, y) + wq1 * dot2v(ang.h[10], x, y) + wq2 * dot2v(ang.h[11], x, y);
        Hh[20] += wq0 * dot2v(ang.h[12], x, y) + wq1 * dot2v(ang.h[13], x, y) + wq2 * dot2v(ang.h[14], x, y);
    }
    {   // w J^T Sigma^-1 J
        const double wxx = w * ixx, wxy = w * ixy, wxz = w * ixz, wyy = w * iyy, wyz = w * iyz, wzz = w * izz;
        double m3[3], m4[3], m5[3];
        m3[0] = wxy * J[0] + wxz * J[1];
        m3[1] = wyy * J[0] + wyz * J[1];
        m3[2] = wyz * J[0] + wzz * J[1];
        m4[0] = wxx * J[2] + wxy * J[3] + wxz * J[4];
        m4[1] = wxy * J[2] + wyy * J[3] + wyz * J[4];
        m4[2] = wxz * J[2] + wyz * J[3] + wzz * J[4];
        m5[0] = wxx * J[5] + wxy * J[6] + wxz * J[7];
        m5[1] = wxy * J[5] + wyy * J[6] + wyz * J[7];
        m5[2] = wxz * J[5] + wyz * J[6] + wzz * J[7];
        Hh[0] += wxx; Hh[1] += wxy; Hh[2] += wxz; Hh[3] += m3[0]; Hh[4] += m4[0]; Hh[5] += m5[0];
        Hh[6] += wyy; Hh[7] += wyz; Hh[8] += m3[1]; Hh[9] += m4[1]; Hh[10] += m5[1];
        Hh[11] += wzz; Hh[12] += m3[2]; Hh[13] += m4[2]; Hh[14] += m5[2];
        Hh[15] += J[0] * m3[1] + J[1] * m3[2];
        Hh[16] += J[0] * m4[1] + J[1] * m4[2];
        Hh[17] += J[0] * m5[1] + J[1] * m5[2];
        Hh[18] += J[2] * m4[0] + J[3] * m4[1] + J[4] * m4[2];
        Hh[19] += J[2] * m5[0] + J[3] * m5[1] + J[4] * m5[2];
        Hh[20] += J[5] * m5[0] + J[6] * m5[1] + J[7] * m5[2];
    }
}

// warp 0 completes a pass request: the twelve sin/cos evaluations run on twelve lanes
__device__ __forceinline__ void finish_request_warp0(NdtSmem &S, int lane) {
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
        if (lane == 0) ctl_finish_request(S.ctl, &S.trig_f[0], &S.trig_f[3], &S.trig_d[0], &S.trig_d[3]);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(NDT_THREADS, NDT_MIN_CTAS) ndt_match_kernel(GridView G, NdtConst K, MatchArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NdtSmem &S = *reinterpret_cast<NdtSmem *>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned C = cluster.num_blocks();
    const unsigned crank = cluster.block_rank();
    const unsigned match = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_compute = warp < NDT_NCW;
    const uint32_t lt = (1u << lane) - 1u;

    uint32_t first = 0, last = A.n_shared;
    if (A.offsets) { first = A.offsets[match]; last = A.offsets[match + 1]; }
    const uint32_t npts = last - first;

    if (tid < NDT_NSW) { S.tail[tid] = 0u; S.head[tid] = 0u; S.finished[tid] = 0u; }
    if (warp == 0) {
        if (lane == 0) {
            if (A.deriv_only) {
                const double *p = A.poses6 + (size_t)match * 6;
                for (int i = 0; i < 6; ++i) S.ctl.p[i] = S.ctl.x_t[i] = p[i];
                ctl_request(S.ctl, p, 1, ST_INIT);
                S.ctl.passes = 0; S.ctl.pairs = 0;
            } else {
                ctl_start(S.ctl, K, A.guesses + (size_t)match * 16, (double)npts);
            }
            S.go = 1;
        }
        finish_request_warp0(S, lane);
    }
    __syncthreads();

    const uint32_t stride = C * NDT_NSW * 32u;
    // The two roles never share code after this point (ptxas sizes each branch for its own register
    // budget); they meet at CTA-wide barriers issued from both branches.
    if (!is_compute) {
        // =========================== search warps (producers) ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(NDT_REG_SEARCH));
        const int sw = warp - NDT_NCW;
        uint2 *ring = S.ring[sw];
        uint32_t pass_id = 0;
        uint32_t my_tail = 0;                       // entries produced so far (warp-uniform)
        while (true) {
            ++pass_id;
            const float *T = S.ctl.T;
            for (uint32_t base = first + (crank * NDT_NSW + sw) * 32u; base < last; base += stride) {
                const uint32_t i = base + lane;
                int ex0 = -1, ex1 = -1, ex2 = -1;
                size_t wbase = 0;
                float tx = 0.f, ty = 0.f, tz = 0.f;
                if (i < last && G.ok) {
                    const float4 pt = __ldg(&A.src[i]);
                    transform_f32(T, pt.x, pt.y, pt.z, tx, ty, tz);
                    if (finite3(tx, ty, tz)) {
                        // every cell that can hold a centroid within the radius (centroids may sit up to
                        // G.margin outside their own cell; the slack also covers the rounding of this arithmetic)
                        const float q[3] = {tx, ty, tz};
                        int lo[3], ex[3];
                        bool empty = false;
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const float mg = G.margin + 1e-6f * fabsf(q[a]);
                            int l = (int)floorf((q[a] - G.res - mg) * G.inv_leaf) - G.min_b[a];
                            int h = (int)floorf((q[a] + G.res + mg) * G.inv_leaf) - G.min_b[a];
                            l = max(l, 0); h = min(h, G.div_b[a] - 1);
                            lo[a] = l; ex[a] = min(h - l, 3);      // the window never exceeds 4 cells (margin << cell)
                            empty = empty || (h < l);
                        }
                        if (!empty) {
                            ex0 = ex[0]; ex1 = ex[1]; ex2 = ex[2];
                            wbase = (size_t)lo[0] + (size_t)lo[1] * G.mul[1] + (size_t)lo[2] * G.mul[2];
                        }
                    }
                }
                const int nplanes = __reduce_max_sync(0xffffffffu, ex2 + 1);
                const bool wide = __any_sync(0xffffffffu, (ex0 > 2) || (ex1 > 2));
                for (int plane = 0; plane < nplanes; ++plane) {
                    // a plane appends at most 32 x 16 entries: wait until the consumer has made room
                    if (lane == 0) {
                        uint32_t hd = ld_vol(&S.head[sw]);
                        while (my_tail - hd > RING - 512u) { __nanosleep(64); hd = ld_vol(&S.head[sw]); }
                    }
                    __syncwarp();
                    const bool act = plane <= ex2;
                    const float4 *wp = G.cells + wbase + (size_t)plane * G.mul[2];
                    uint32_t qn = my_tail;
                    if (!wide) {
                        float4 c[9];
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) {
                                c[dy * 3 + dx] = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (act && dx <= ex0 && dy <= ex1) c[dy * 3 + dx] = __ldg(wp + dx + (size_t)dy * G.mul[1]);
                            }
#pragma unroll
                        for (int k = 0; k < 9; ++k) {
                            // flann::L2_Simple<float>: (dx*dx + dy*dy) + dz*dz, accepted when < (float)(r*r)
                            const float dx = __fsub_rn(tx, c[k].x), dy = __fsub_rn(ty, c[k].y), dz = __fsub_rn(tz, c[k].z);
                            const float d2f = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                            const int code = __float_as_int(c[k].w);
                            const bool hit = (code > 0) && (d2f < G.r2);
                            const uint32_t b = __ballot_sync(0xffffffffu, hit);
                            if (hit) ring[(qn + __popc(b & lt)) & (RING - 1u)] = make_uint2(i, (uint32_t)(code - 1));
                            qn += __popc(b);
                        }
                    } else {
                        // some lane's window is 4 cells wide (query within `margin` of a cell face): rare
                        for (int dy = 0; dy < 4; ++dy)
                            for (int dx = 0; dx < 4; ++dx) {
                                float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (act && dx <= ex0 && dy <= ex1) c = __ldg(wp + dx + (size_t)dy * G.mul[1]);
                                const float ddx = __fsub_rn(tx, c.x), ddy = __fsub_rn(ty, c.y), ddz = __fsub_rn(tz, c.z);
                                const float d2f = __fadd_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)), __fmul_rn(ddz, ddz));
                                const int code = __float_as_int(c.w);
                                const bool hit = (code > 0) && (d2f < G.r2);
                                const uint32_t b = __ballot_sync(0xffffffffu, hit);
                                if (hit) ring[(qn + __popc(b & lt)) & (RING - 1u)] = make_uint2(i, (uint32_t)(code - 1));
                                qn += __popc(b);
                            }
                    }
                    __syncwarp();
                    if (qn != my_tail) {
                        my_tail = qn;
                        if (lane == 0) { __threadfence_block(); st_vol(&S.tail[sw], my_tail); }
                    }
                }
            }
            // publish the end of this warp's stream for this pass
            __syncwarp();
            if (lane == 0) { __threadfence_block(); st_vol(&S.tail[sw], my_tail); __threadfence_block(); st_vol(&S.finished[sw], pass_id); }
            cta_barrier();                          // (1) all pairs of the pass consumed, partials written
            if (C > 1) cluster.sync();
            cta_barrier();                          // (2) totals ready
            cta_barrier();                          // (3) controller done
            if (!ld_vol(reinterpret_cast<const uint32_t *>(&S.go))) break;
        }
        if (C > 1) cluster.sync();
    } else {
    // =========================== compute warps (consumers) ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(NDT_REG_COMPUTE));
    int parity = 0;
    uint32_t pass_id = 0;
    uint32_t cpos[NDT_PPC];                         // entries consumed per producer
#pragma unroll
    for (int k = 0; k < NDT_PPC; ++k) cpos[k] = 0;
    while (true) {
        ++pass_id;
        {
            const float *T = S.ctl.T;
            const bool hess = S.ctl.hess != 0;
            const AngTab &ang = S.ctl.ang;
            double acc[ACC_N];
#pragma unroll
            for (int i = 0; i < ACC_N; ++i) acc[i] = 0.0;
            uint32_t done_mask = 0;                         // bit k: producer k exhausted for this pass
            int turn = 0;
            while (done_mask != (1u << NDT_PPC) - 1u) {
                // fixed round-robin over this warp's producers (deterministic accumulation order)
                const int k = turn;
                turn = (turn + 1 == NDT_PPC) ? 0 : turn + 1;
                if (done_mask & (1u << k)) continue;
                const int sw = warp + k * NDT_NCW;
                uint32_t pos = 0;
#pragma unroll
                for (int kk = 0; kk < NDT_PPC; ++kk) if (kk == k) pos = cpos[kk];
                // wait for a full chunk of 32 pairs, or for the producer to finish the pass
                uint32_t n = 0;
                if (lane == 0) {
                    while (true) {
                        const uint32_t fin = ld_vol(&S.finished[sw]);
                        __threadfence_block();
                        const uint32_t avail = ld_vol(&S.tail[sw]) - pos;
                        if (avail >= 32u) { n = 32u; break; }
                        if (fin == pass_id) { n = avail | 0x80000000u; break; }     // final (possibly empty) chunk
                        __nanosleep(32);
                    }
                }
                n = __shfl_sync(0xffffffffu, n, 0);
                const bool final_chunk = (n & 0x80000000u) != 0;
                n &= 0x7fffffffu;
                __threadfence_block();
                if ((uint32_t)lane < n) {
                    const uint2 e = S.ring[sw][(pos + lane) & (RING - 1u)];
                    const float4 sp = __ldg(&A.src[e.x]);
                    ndt_pair(sp.x, sp.y, sp.z, T, ang, G.gauss + (size_t)e.y * 10, K.d1, K.d2, hess, acc);
                }
                pos += n;
#pragma unroll
                for (int kk = 0; kk < NDT_PPC; ++kk) if (kk == k) cpos[kk] = pos;
                __syncwarp();
                if (lane == 0 && n) st_vol(&S.head[sw], pos);
                if (final_chunk) done_mask |= (1u << k);
            }
            // warp butterfly (fixed order) -> one partial per compute warp
#pragma unroll
            for (int i = 0; i < ACC_N; ++i) {
                double vsum = acc[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
                if (lane == 0) S.warp_part[warp][i] = vsum;
            }
        }
        cta_barrier();                              // (1)
        // ---------------- deterministic reduction: CTA -> cluster (fixed order) ----------------
        if (tid < ACC_N) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NDT_NCW; ++w) s += S.warp_part[w][tid];
            S.cta_part[parity][tid] = s;
        }
        if (C > 1) {
            cluster.sync();
            if (tid < ACC_N) {
                double tot = 0.0;
                for (unsigned r = 0; r < C; ++r) tot += *cluster.map_shared_rank(&S.cta_part[parity][tid], r);
                S.total[tid] = tot;
            }
        } else {
            if (tid < ACC_N) S.total[tid] = S.cta_part[parity][tid];
        }
        cta_barrier();                              // (2)
        // ---------------- controller: Newton step + More-Thuente state machine (warp 0) ----------------
        if (warp == 0) {
            int go = 0;
            if (lane == 0) { go = A.deriv_only ? 0 : ctl_step(S.ctl, K, S.total); S.go = go; }
            go = __shfl_sync(0xffffffffu, go, 0);
            if (go) finish_request_warp0(S, lane);
        }
        cta_barrier();                              // (3)
        if (!S.go) break;
        parity ^= 1;
    }
    if (C > 1) cluster.sync();    // peers may still be reading this CTA's partials
    if (tid == 0 && crank == 0) {
        if (A.deriv_only) {
            for (int i = 0; i < ACC_N; ++i) A.acc_out[(size_t)match * ACC_N + i] = S.total[i];
            if (A.poses_out) for (int i = 0; i < 16; ++i) A.poses_out[(size_t)match * 16 + i] = S.ctl.T[i];
        } else {
            for (int i = 0; i < 16; ++i) A.poses_out[(size_t)match * 16 + i] = S.ctl.finalT[i];
            if (A.results) {
                b2ndt_result r;
                r.iterations = S.ctl.nr_iter; r.converged = S.ctl.converged;
                r.score = S.ctl.score; r.trans_probability = S.ctl.trans_probability;
                for (int i = 0; i < 6; ++i) r.p[i] = S.ctl.p[i];
                r.passes = S.ctl.passes; r.mt_trials = S.ctl.mt_trials; r.pairs = S.ctl.pairs;
                A.results[match] = r;
            }
        }
    }
    }   // compute branch
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
constexpr int FIT_RMAX = 8;

__device__ __forceinline__ void fit_cell(const FitView &F, int kx, int ky, int kz, float qx, float qy, float qz, float &best) {
    const int32_t v = __float_as_int(__ldg(F.cells + (size_t)kx + (size_t)ky * F.mul[1] + (size_t)kz * F.mul[2]).w);
    if (v == 0) return;
    const int j = (v > 0 ? v : -v) - 1;
    const uint32_t s = __ldg(&F.leaf_start[j]);
    const uint32_t n = (uint32_t)__ldg(&F.leaf_n[j]);
    for (uint32_t k = s; k < s + n; ++k) {
        const float4 p = __ldg(&F.pts_sorted[k]);
        const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        best = fminf(best, d2);
    }
}

__global__ void __launch_bounds__(FIT_THREADS) fitness_kernel(FitView F, const float4 *__restrict__ src, uint32_t n, PoseArg P,
                                                              double max_range, double *__restrict__ part_sum,
                                                              unsigned long long *__restrict__ part_cnt) {
    double sum = 0.0;
    unsigned long long cnt = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 pt = __ldg(&src[i]);
        float qx, qy, qz;
        transform_f32(P.T, pt.x, pt.y, pt.z, qx, qy, qz);
        if (!F.ok || !finite3(qx, qy, qz)) continue;
        const int cx = (int)floorf(qx * F.inv_leaf) - F.min_b[0];
        const int cy = (int)floorf(qy * F.inv_leaf) - F.min_b[1];
        const int cz = (int)floorf(qz * F.inv_leaf) - F.min_b[2];
        float best = FLT_MAX;
        bool done = false;
        for (int r = 0; r <= FIT_RMAX && !done; ++r) {
            const int z0 = max(cz - r, 0), z1 = min(cz + r, F.div_b[2] - 1);
            const int y0 = max(cy - r, 0), y1 = min(cy + r, F.div_b[1] - 1);
            const int x0 = max(cx - r, 0), x1 = min(cx + r, F.div_b[0] - 1);
            for (int kz = z0; kz <= z1; ++kz)
                for (int ky = y0; ky <= y1; ++ky) {
                    const bool face = (abs(kz - cz) == r) || (abs(ky - cy) == r);
                    if (face) {
                        for (int kx = x0; kx <= x1; ++kx) fit_cell(F, kx, ky, kz, qx, qy, qz, best);
                    } else {
                        if (cx - r >= 0 && cx - r < F.div_b[0]) fit_cell(F, cx - r, ky, kz, qx, qy, qz, best);
                        if (r > 0 && cx + r >= 0 && cx + r < F.div_b[0]) fit_cell(F, cx + r, ky, kz, qx, qy, qz, best);
                    }
                }
            // everything not yet visited lies at least r cells away from the query
            const double lim = (double)r * (double)F.res * (1.0 - 1e-4);
            if (best < FLT_MAX && (double)best * (1.0 + 1e-5) <= lim * lim) done = true;
            // the rings already cover the whole grid
            if (cx - r <= 0 && cy - r <= 0 && cz - r <= 0 && cx + r >= F.div_b[0] - 1 && cy + r >= F.div_b[1] - 1 &&
                cz + r >= F.div_b[2] - 1)
                done = true;
        }
        if (!done) {   // rare: isolated query, exact brute force over all target points
            for (uint32_t k = 0; k < F.N; ++k) {
                const float4 p = __ldg(&F.pts_sorted[k]);
                const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
                best = fminf(best, __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
            }
        }
        if (best < FLT_MAX && (double)best <= max_range) { sum += (double)best; ++cnt; }
    }
    __shared__ double ssum[FIT_THREADS / 32];
    __shared__ unsigned long long scnt[FIT_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
    if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = sum; scnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0; unsigned long long c = 0;
        for (int w = 0; w < FIT_THREADS / 32; ++w) { s += ssum[w]; c += scnt[w]; }
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
    float last_pose[16];
    bool have_last = false;
    int cl_single = 8, cl_batch = 1;
    bool attrs_set = false;
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
    h->d_src.release(); h->d_guess.release(); h->d_pose.release(); h->d_res.release(); h->d_off.release();
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
    if (single_match_ctas < 1 || single_match_ctas > 16 || batch_ctas < 1 || batch_ctas > 16) {
        set_error("b2ndt_set_cluster: cluster width must be 1..16"); return B2_ERR_INVALID;
    }
    h->cl_single = single_match_ctas; h->cl_batch = batch_ctas;
    return 0;
}

static int build_target(b2ndt *h, const float4 *d_pts, size_t n, int nbits_hint) {
    TargetDev &t = h->tgt;
    t.valid = false; t.N = (uint32_t)n; t.V = 0; t.n_tree = 0; t.max_disp = 0.f;
    h->have_last = false;
    memset(&t.L, 0, sizeof(t.L));
    int rc;
    uint32_t off[2] = {0u, (uint32_t)n};
    if ((rc = h->pipe.plan(off, 1, h->st))) return rc;
    const float res = h->prm.res;
    if ((rc = h->pipe.run(d_pts, res, res, res, nbits_hint, h->st))) return rc;
    if ((rc = h->h_small.reserve(4096))) return rc;
    uint32_t *misc = h->h_small.as<uint32_t>();
    B2_CUDA(cudaMemcpyAsync(misc, h->pipe.scalars(), 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaMemcpyAsync(misc + 8, h->pipe.layouts(), sizeof(VoxLayout), cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaStreamSynchronize(h->st));
    memcpy(&t.L, misc + 8, sizeof(VoxLayout));
    t.valid = true;
    if (!t.L.ok) return 0;     // empty cloud or PCL's int32 guard: no cells (PCL warns and clears)
    t.V = misc[1];
    const uint32_t V = t.V;
    if ((rc = t.pts_sorted.reserve((n + 1) * sizeof(float4)))) return rc;
    if ((rc = t.leaf_idx.reserve((V + 1) * 4))) return rc;
    if ((rc = t.leaf_n.reserve((V + 1) * 4))) return rc;
    if ((rc = t.leaf_start.reserve((V + 1) * 4))) return rc;
    if ((rc = t.centroid4.reserve((V + 1) * sizeof(float4)))) return rc;
    if ((rc = t.gauss.reserve((size_t)(V + 1) * 80))) return rc;
    if ((rc = t.sums.reserve((size_t)(V + 1) * 72))) return rc;
    if ((rc = t.icov9.reserve((size_t)(V + 1) * 72))) return rc;
    if ((rc = t.cells.reserve((size_t)t.L.ncells * 16 + 16))) return rc;
    if ((rc = t.counters.reserve(64))) return rc;
    B2_CUDA(cudaMemsetAsync(t.cells.p, 0, (size_t)t.L.ncells * 16, h->st));
    B2_CUDA(cudaMemsetAsync(t.counters.p, 0, 64, h->st));
    if (V) {
        unsigned blocks = (V + 7) / 8;
        if (blocks > 148 * 32) blocks = 148 * 32;
        leaf_stats_kernel<<<blocks, 256, 0, h->st>>>(d_pts, h->pipe.sorted_keys(), h->pipe.sorted_vals(), h->pipe.run_start(), V,
                                                    t.L.n_finite, t.pts_sorted.as<float4>(), t.leaf_idx.as<int32_t>(),
                                                    t.leaf_n.as<int32_t>(), t.leaf_start.as<uint32_t>(), t.centroid4.as<float4>(),
                                                    t.sums.as<double>());
        B2_LAUNCH_CHECK();
        LayoutArg LA;
        for (int a = 0; a < 3; ++a) { LA.min_b[a] = t.L.min_b[a]; LA.div_b[a] = t.L.div_b[a]; LA.inv[a] = t.L.inv[a]; }
        leaf_finish_kernel<<<(V + 127) / 128, 128, 0, h->st>>>(V, h->prm.min_pts, h->prm.eig_mult, LA, t.leaf_idx.as<int32_t>(),
                                                              t.leaf_n.as<int32_t>(), t.centroid4.as<float4>(), t.sums.as<double>(),
                                                              t.gauss.as<double>(), t.icov9.as<double>(), t.cells.as<float4>(),
                                                              t.counters.as<uint32_t>());
        B2_LAUNCH_CHECK();
    }
    B2_CUDA(cudaMemcpyAsync(misc, t.counters.p, 8, cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaStreamSynchronize(h->st));
    t.n_tree = misc[0];
    memcpy(&t.max_disp, &misc[1], 4);
    return 0;
}

static int check_cloud_args(const char *fn, const void *pts, size_t n, size_t stride, size_t ioff) {
    if (n && !pts) { set_error("%s: NULL cloud", fn); return B2_ERR_INVALID; }
    if (stride < 16 || (stride & 3) || ioff + 4 > stride || (ioff & 3)) { set_error("%s: bad stride %zu / intensity offset %zu", fn, stride, ioff); return B2_ERR_INVALID; }
    if (n >= 0xFFFFFFF0ull) { set_error("%s: cloud too large", fn); return B2_ERR_INVALID; }
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
    if (n >= 0xFFFFFFF0ull) { set_error("b2ndt_set_target_device: cloud too large"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(h->device));
    return build_target(h, (const float4 *)d_pts_f4, n, 0);
}

extern "C" int b2ndt_target_info_get(b2ndt *h, b2ndt_target_info *info) {
    if (!h || !info) { set_error("b2ndt_target_info_get: NULL argument"); return B2_ERR_INVALID; }
    if (!h->tgt.valid) { set_error("b2ndt_target_info_get: no target set"); return B2_ERR_STATE; }
    const TargetDev &t = h->tgt;
    info->ok = t.L.ok;
    for (int a = 0; a < 3; ++a) { info->min_b[a] = t.L.min_b[a]; info->div_b[a] = t.L.div_b[a]; }
    info->n_points = t.L.n_finite; info->n_leaves = t.V; info->n_tree = t.n_tree; info->inv_leaf = t.L.inv[0];
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
        std::vector<double> g(V * 10);
        B2_CUDA(cudaMemcpy(g.data(), t.gauss.p, V * 80, cudaMemcpyDeviceToHost));
        for (size_t j = 0; j < V; ++j) { mean3[3 * j] = g[10 * j]; mean3[3 * j + 1] = g[10 * j + 1]; mean3[3 * j + 2] = g[10 * j + 2]; }
    }
    return 0;
}

static GridView make_grid_view(const b2ndt *h) {
    GridView G;
    memset(&G, 0, sizeof(G));
    const TargetDev &t = h->tgt;
    G.cells = t.cells.as<float4>();
    G.gauss = t.gauss.as<double>();
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
        B2_CUDA(cudaFuncSetAttribute(ndt_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NdtSmem)));
        B2_CUDA(cudaFuncSetAttribute(ndt_match_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        h->attrs_set = true;
    }
    GridView G = make_grid_view(h);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(B * (size_t)C));
    cfg.blockDim = dim3(NDT_THREADS);
    cfg.dynamicSmemBytes = sizeof(NdtSmem);
    cfg.stream = h->st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    NdtConst K = h->K;
    MatchArgs Ac = A;
    cudaError_t e = cudaLaunchKernelEx(&cfg, ndt_match_kernel, G, K, Ac);
    b2::count_launch();
    if (e != cudaSuccess) { set_error("ndt_match_kernel launch failed: %s", cudaGetErrorString(e)); return B2_ERR_CUDA; }
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
    B2_CUDA(cudaMemcpyAsync(h->d_guess.p, gst, B * 64, cudaMemcpyHostToDevice, h->st));
    const uint32_t *d_off = nullptr;
    if (offsets) {
        if ((rc = h->d_off.reserve((B + 1) * 4))) return rc;
        uint32_t *ost = (uint32_t *)(stage + src_stage + B * 64);
        memcpy(ost, offsets, (B + 1) * 4);
        B2_CUDA(cudaMemcpyAsync(h->d_off.p, ost, (B + 1) * 4, cudaMemcpyHostToDevice, h->st));
        d_off = h->d_off.as<uint32_t>();
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
        if ((rc = launch_match(h, A, c1 - c0, C))) return rc;
    }
    char *hres = h->h_res.as<char>();
    B2_CUDA(cudaMemcpyAsync(hres, h->d_pose.p, B * 64, cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaMemcpyAsync(hres + B * 64, h->d_res.p, B * sizeof(b2ndt_result), cudaMemcpyDeviceToHost, h->st));
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
    int C = h->cl_single;
    while (C > 1 && (size_t)C * NDT_NSW * 32 > n) C >>= 1;
    rc = align_host(h, src, n, stride, ioff, nullptr, 1, guess, pose_out, res, C < 1 ? 1 : C);
    if (rc) return rc;
    h->last_n = n;
    memcpy(h->last_pose, pose_out, 64);
    h->have_last = true;
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
    memcpy(stage + n * 16, p, 48);
    B2_CUDA(cudaMemcpyAsync(h->d_p6.p, stage + n * 16, 48, cudaMemcpyHostToDevice, h->st));
    MatchArgs A;
    memset(&A, 0, sizeof(A));
    A.src = h->d_src.as<float4>(); A.n_shared = (uint32_t)n; A.poses6 = h->d_p6.as<double>();
    A.acc_out = h->d_acc.as<double>(); A.deriv_only = 1;
    int C = h->cl_single;
    while (C > 1 && (size_t)C * NDT_NSW * 32 > n) C >>= 1;
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
    unsigned blocks = (unsigned)((n + FIT_THREADS - 1) / FIT_THREADS);
    if (blocks > 148 * 8) blocks = 148 * 8;
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
    return fitness_device(h, h->d_src.as<float4>(), h->last_n, h->last_pose, max_range, out);
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
