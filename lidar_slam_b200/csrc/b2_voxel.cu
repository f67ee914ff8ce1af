// b2_voxel.cu -- voxelisation pipeline kernels (bbox, PCL voxel layout, keys, batched stable LSD radix
// sort, run heads) and the VoxelFilter C ABI (b2vf_*).
//
// Reference semantics: pcl::VoxelGrid<PointXYZI>::applyFilter as called from
// lidar_localization/src/models/cloud_filter/voxel_filter.cpp:36-41.
// All of this is HBM-bound integer / byte work: coalesced 16-byte point loads, one CTA per 4096-key
// tile, no tensor cores.
#include "b2_voxel.cuh"
#include "b2_cloud.cuh"

#include <float.h>
#include <stdarg.h>
#include <stdlib.h>

#include <cmath>

#include <thread>

namespace b2 {

// ------------------------------------------------------------------ error plumbing ----------
static thread_local std::string t_err;
std::atomic<uint64_t> g_launches{0};
void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_err = buf;
}
const char *last_error() { return t_err.c_str(); }

// Upper bound on the voxel-key width from a host-side bounding box (PCL layout arithmetic, one extra
// cell per axis of slack); lets the sort run only the radix passes it needs without a device round trip.
int key_bits_from_bbox(const float mn[3], const float mx[3], float lx, float ly, float lz) {
    const float leaf[3] = {lx, ly, lz};
    double cells = 1.0;
    for (int a = 0; a < 3; ++a) {
        if (!(mx[a] >= mn[a])) return 0;
        const float inv = 1.0f / leaf[a];
        double d = floor((double)mx[a] * inv) - floor((double)mn[a] * inv) + 2.0;
        cells *= d;
    }
    if (!(cells < 2147483000.0)) return 32;
    uint32_t c = (uint32_t)cells;
    int bits = 0;
    while (c) { ++bits; c >>= 1; }
    return bits < 1 ? 1 : bits;
}

void pack_cloud_f4_bbox(const void *src, size_t n, size_t stride, size_t ioff, float *dst, float mn[3], float mx[3]) {
    struct BB { float mn[3], mx[3]; };
    auto work = [=](size_t a, size_t b, BB *bb) {
        const char *s = (const char *)src;
        float lmn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, lmx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (size_t i = a; i < b; ++i) {
            const float *p = (const float *)(s + i * stride);
            const float x = p[0], y = p[1], z = p[2];
            dst[4 * i + 0] = x; dst[4 * i + 1] = y; dst[4 * i + 2] = z; dst[4 * i + 3] = *(const float *)(s + i * stride + ioff);
            if (std::isfinite(x) && std::isfinite(y) && std::isfinite(z)) {
                lmn[0] = x < lmn[0] ? x : lmn[0]; lmn[1] = y < lmn[1] ? y : lmn[1]; lmn[2] = z < lmn[2] ? z : lmn[2];
                lmx[0] = x > lmx[0] ? x : lmx[0]; lmx[1] = y > lmx[1] ? y : lmx[1]; lmx[2] = z > lmx[2] ? z : lmx[2];
            }
        }
        for (int k = 0; k < 3; ++k) { bb->mn[k] = lmn[k]; bb->mx[k] = lmx[k]; }
    };
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = n > 400000 ? (hw > 8 ? 8 : (hw ? hw : 1)) : 1;
    BB bbs[8];
    for (size_t t = 0; t < 8; ++t) for (int k = 0; k < 3; ++k) { bbs[t].mn[k] = FLT_MAX; bbs[t].mx[k] = -FLT_MAX; }
    if (nt <= 1) work(0, n, &bbs[0]);
    else {
        // no C++ exception may cross the C ABI: if a worker thread cannot be created, its share is done here
        std::thread th[8];
        size_t chunk = (n + nt - 1) / nt;
        for (size_t t = 0; t < nt; ++t) {
            size_t a = t * chunk, b = a + chunk < n ? a + chunk : n;
            if (a >= b) continue;
            try { th[t] = std::thread(work, a, b, &bbs[t]); } catch (...) { work(a, b, &bbs[t]); }
        }
        for (size_t t = 0; t < nt; ++t) if (th[t].joinable()) th[t].join();
    }
    for (int k = 0; k < 3; ++k) { mn[k] = FLT_MAX; mx[k] = -FLT_MAX; }
    for (size_t t = 0; t < nt; ++t)
        for (int k = 0; k < 3; ++k) { mn[k] = bbs[t].mn[k] < mn[k] ? bbs[t].mn[k] : mn[k]; mx[k] = bbs[t].mx[k] > mx[k] ? bbs[t].mx[k] : mx[k]; }
}

void pack_cloud_f4(const void *src, size_t n, size_t stride, size_t ioff, float *dst) {
    auto work = [=](size_t a, size_t b) {
        const char *s = (const char *)src;
        for (size_t i = a; i < b; ++i) {
            const float *p = (const float *)(s + i * stride);
            float inten = *(const float *)(s + i * stride + ioff);
            dst[4 * i + 0] = p[0]; dst[4 * i + 1] = p[1]; dst[4 * i + 2] = p[2]; dst[4 * i + 3] = inten;
        }
    };
    if (stride == 16 && ioff == 12 && n <= 400000) { memcpy(dst, src, n * 16); return; }
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = n > 400000 ? (hw > 8 ? 8 : (hw ? hw : 1)) : 1;
    if (nt <= 1) { work(0, n); return; }
    std::thread th[8];
    size_t chunk = (n + nt - 1) / nt;
    for (size_t t = 0; t < nt; ++t) {
        size_t a = t * chunk, b = a + chunk < n ? a + chunk : n;
        if (a >= b) continue;
        try { th[t] = std::thread(work, a, b); } catch (...) { work(a, b); }     // see pack_cloud_f4_bbox
    }
    for (size_t t = 0; t < nt; ++t) if (th[t].joinable()) th[t].join();
}

// ------------------------------------------------------------------ kernels -----------------
// Look-back words: two flag bits + a 30-bit count (B2_MAX_POINTS < 2^30).
constexpr uint32_t LB_PART = 1u << 30, LB_INCL = 2u << 30, LB_MASK = 0x3FFFFFFFu;
__device__ __forceinline__ uint32_t ld_vol_g(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void st_vol_g(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }

// exclusive scan of one value per thread over a 256-thread CTA; *total = the CTA's sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *wsum /*smem[9]*/, uint32_t *total) {
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o); if (l >= o) inc += n; }
    if (l == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        const uint32_t t = (l < SORT_THREADS / 32) ? wsum[l] : 0u;
        uint32_t ti = t;
#pragma unroll
        for (int o = 1; o < SORT_THREADS / 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, ti, o); if (l >= o) ti += n; }
        if (l < SORT_THREADS / 32) wsum[l] = ti - t;
        if (l == SORT_THREADS / 32 - 1) wsum[SORT_THREADS / 32] = ti;
    }
    __syncthreads();
    const uint32_t r = wsum[w] + inc - v;
    if (total) *total = wsum[SORT_THREADS / 32];
    __syncthreads();
    return r;
}

// pcl::VoxelGrid::applyFilter prologue: inverse leaf, int64 overflow guard, min_b / div_b / divb_mul
__device__ void compute_layout(const float mn[3], const float mx[3], uint32_t n_finite, float lx, float ly, float lz, VoxLayout &L) {
    memset(&L, 0, sizeof(L));
    const float leaf[3] = {lx, ly, lz};
    for (int a = 0; a < 3; ++a) L.inv[a] = __fdiv_rn(1.0f, leaf[a]);
    L.n_finite = n_finite;
    if (L.n_finite > 0) {
        long long d[3];
        for (int a = 0; a < 3; ++a) {
            L.min_p[a] = mn[a];
            L.max_p[a] = mx[a];
            d[a] = (long long)__fmul_rn(__fsub_rn(L.max_p[a], L.min_p[a]), L.inv[a]) + 1;
        }
        bool ok = (d[0] * d[1] * d[2]) <= 2147483647LL;
        if (ok) {
            unsigned long long nc = 1;
            for (int a = 0; a < 3; ++a) {
                L.min_b[a] = (int)floorf(__fmul_rn(L.min_p[a], L.inv[a]));
                int mxb = (int)floorf(__fmul_rn(L.max_p[a], L.inv[a]));
                L.div_b[a] = mxb - L.min_b[a] + 1;
                nc *= (unsigned long long)L.div_b[a];
            }
            L.mul[0] = 1; L.mul[1] = L.div_b[0]; L.mul[2] = L.div_b[0] * L.div_b[1];
            if (nc > 0x7FFFFFF0ull) ok = false;   // div_b can exceed the guard's estimate by one per axis
            L.ncells = (uint32_t)nc;
        }
        L.ok = ok ? 1 : 0;
    }
    if (L.ok) L.nbits = 32 - __clz(L.ncells);   // keys lie in [0, ncells] (ncells = non-finite points)
}

// ---- kernel 1: per-tile bounding boxes; the LAST CTA to finish (ticket in scalars[2]) folds them per cloud, writes
// the PCL layouts and resets the tickets of the kernels that follow.  Also clears the digit histograms.
// Op is the fused ingest (b2_cloud.cuh): the point is transformed / dropped first, the result (a dropped point = a NaN point)
// is written to pts_out for the kernels that follow, and the box is taken over the results.
template <class Op>
__global__ void __launch_bounds__(SORT_THREADS) bbox_layout_kernel(const float4 *__restrict__ pts, float4 *__restrict__ pts_out, Op op,
                                                                   PlanView P, float lx, float ly,
                                                                   float lz, float *__restrict__ part,
                                                                   VoxLayout *__restrict__ layouts, uint32_t *__restrict__ scalars,
                                                                   uint32_t *__restrict__ hist, uint32_t hist_words) {
    chain_sync();
    const TileDesc t = get_tile(P, blockIdx.x);
    float mn0 = FLT_MAX, mn1 = FLT_MAX, mn2 = FLT_MAX, mx0 = -FLT_MAX, mx1 = -FLT_MAX, mx2 = -FLT_MAX;
    uint32_t cnt = 0;
    const uint32_t seg_begin = get_seg(P, t.seg).begin;
    for (uint32_t k = threadIdx.x; k < t.count; k += SORT_THREADS) {
        float4 p;
        const bool keep = op(__ldg(&pts[t.begin + k]), t.begin + k - seg_begin, p);
        if (pts_out) pts_out[t.begin + k] = p;
        if (keep) {
            mn0 = fminf(mn0, p.x); mn1 = fminf(mn1, p.y); mn2 = fminf(mn2, p.z);
            mx0 = fmaxf(mx0, p.x); mx1 = fmaxf(mx1, p.y); mx2 = fmaxf(mx2, p.z);
            ++cnt;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn0 = fminf(mn0, __shfl_xor_sync(0xffffffffu, mn0, o));
        mn1 = fminf(mn1, __shfl_xor_sync(0xffffffffu, mn1, o));
        mn2 = fminf(mn2, __shfl_xor_sync(0xffffffffu, mn2, o));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
        mx2 = fmaxf(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    __shared__ float s[SORT_THREADS / 32][6];
    __shared__ uint32_t sc[SORT_THREADS / 32];
    __shared__ uint32_t s_last, s_bits;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s[w][0] = mn0; s[w][1] = mn1; s[w][2] = mn2; s[w][3] = mx0; s[w][4] = mx1; s[w][5] = mx2; sc[w] = cnt; }
    {   // this CTA's slice of the histogram words
        const uint32_t chunk = (hist_words + gridDim.x - 1) / gridDim.x;
        const uint32_t lo = blockIdx.x * chunk, hi = (lo + chunk < hist_words) ? lo + chunk : hist_words;
        for (uint32_t i = lo + threadIdx.x; i < hi; i += SORT_THREADS) hist[i] = 0u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t c = 0;
        for (int i = 0; i < SORT_THREADS / 32; ++i) {
            c += sc[i];
            mn0 = fminf(mn0, s[i][0]); mn1 = fminf(mn1, s[i][1]); mn2 = fminf(mn2, s[i][2]);
            mx0 = fmaxf(mx0, s[i][3]); mx1 = fmaxf(mx1, s[i][4]); mx2 = fmaxf(mx2, s[i][5]);
        }
        float *o = part + (size_t)blockIdx.x * 8;
        o[0] = mn0; o[1] = mn1; o[2] = mn2; o[3] = mx0; o[4] = mx1; o[5] = mx2; o[6] = __uint_as_float(c);
        __threadfence();
        s_last = (atomicAdd(&scalars[2], 1u) == gridDim.x - 1u) ? 1u : 0u;
        s_bits = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (uint32_t sg_i = w; sg_i < P.B; sg_i += SORT_THREADS / 32) {      // one warp per cloud
        const SegDesc sg = get_seg(P, sg_i);
        float a0 = FLT_MAX, a1 = FLT_MAX, a2 = FLT_MAX, b0 = -FLT_MAX, b1 = -FLT_MAX, b2 = -FLT_MAX;
        uint32_t c = 0;
        for (uint32_t k = l; k < sg.ntiles; k += 32) {
            const float *o = part + (size_t)(sg.tile_begin + k) * 8;
            const float4 lo4 = __ldcg(reinterpret_cast<const float4 *>(o)), hi4 = __ldcg(reinterpret_cast<const float4 *>(o) + 1);
            a0 = fminf(a0, lo4.x); a1 = fminf(a1, lo4.y); a2 = fminf(a2, lo4.z);
            b0 = fmaxf(b0, lo4.w); b1 = fmaxf(b1, hi4.x); b2 = fmaxf(b2, hi4.y);
            c += __float_as_uint(hi4.z);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 = fminf(a0, __shfl_xor_sync(0xffffffffu, a0, o)); a1 = fminf(a1, __shfl_xor_sync(0xffffffffu, a1, o));
            a2 = fminf(a2, __shfl_xor_sync(0xffffffffu, a2, o)); b0 = fmaxf(b0, __shfl_xor_sync(0xffffffffu, b0, o));
            b1 = fmaxf(b1, __shfl_xor_sync(0xffffffffu, b1, o)); b2 = fmaxf(b2, __shfl_xor_sync(0xffffffffu, b2, o));
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (l == 0) {
            const float mn[3] = {a0, a1, a2}, mx[3] = {b0, b1, b2};
            VoxLayout L;
            compute_layout(mn, mx, c, lx, ly, lz, L);
            layouts[sg_i] = L;
            if (L.ok) atomicMax(&s_bits, (uint32_t)L.nbits);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        scalars[0] = s_bits; scalars[1] = 0u; scalars[2] = 0u;
        for (int i = 4; i < 12; ++i) scalars[i] = 0u;      // tickets of the kernels that follow + work-list counters
    }
}

// ---- kernel 2: voxel key per point + the digit histograms of EVERY radix pass (one read of the points), and the
// clearing of this tile's look-back words.  Pass 0 digits are diverse (plain shared-memory atomics); the digits of the
// higher passes are mostly equal inside a warp, so they are aggregated with match.any first.
template <bool FROM_POINTS>
__global__ void __launch_bounds__(SORT_THREADS) key_hist_kernel(const float4 *__restrict__ pts, PlanView P,
                                                                const VoxLayout *__restrict__ layouts, uint32_t *__restrict__ keys,
                                                                uint32_t *__restrict__ hist, uint32_t *__restrict__ state,
                                                                const uint32_t *__restrict__ scalars) {
    chain_sync();
    const TileDesc t = get_tile(P, blockIdx.x);
    __shared__ VoxLayout L;
    __shared__ uint32_t h[4][RADIX];
    if (FROM_POINTS && threadIdx.x == 0) L = layouts[t.seg];
    for (int i = threadIdx.x; i < 4 * RADIX; i += SORT_THREADS) (&h[0][0])[i] = 0u;
    __syncthreads();
    const int npass = (int)((scalars[0] + RADIX_BITS - 1) / RADIX_BITS);
    const int l = threadIdx.x & 31;
    for (uint32_t k0 = threadIdx.x & ~31u; k0 < t.count; k0 += SORT_THREADS) {      // warp-uniform trip count
        const uint32_t k = k0 + l;
        const bool valid = k < t.count;
        uint32_t key = 0u;
        if (valid) {
            if (FROM_POINTS) {
                const float4 p = __ldg(&pts[t.begin + k]);
                key = L.ncells;
                if (L.ok && finite3(p.x, p.y, p.z)) key = (uint32_t)vox_index(L, p.x, p.y, p.z);
                keys[t.begin + k] = key;
            } else {
                key = __ldg(&keys[t.begin + k]);
            }
            atomicAdd(&h[0][key & (RADIX - 1)], 1u);
        }
        for (int p = 1; p < npass; ++p) {
            const uint32_t d = valid ? ((key >> (p * RADIX_BITS)) & (RADIX - 1)) : (uint32_t)RADIX;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            if (valid && l == __ffs(peers) - 1) atomicAdd(&h[p][d], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    for (int p = 0; p < npass; ++p) {
        const uint32_t v = h[p][threadIdx.x];
        if (v) atomicAdd(&hist[((size_t)t.seg * 4 + p) * RADIX + threadIdx.x], v);
    }
    for (int p = 0; p < 4; ++p) state[((size_t)p * P.ntiles + blockIdx.x) * RADIX + threadIdx.x] = 0u;
    if (threadIdx.x == 0) state[(size_t)4 * P.ntiles * RADIX + blockIdx.x] = 0u;
}

// ---- kernel 3: ONE radix pass in ONE kernel ("onesweep"): every tile ranks its keys (match.any + per-warp digit
// counters, stable by construction), publishes its digit counts, obtains the counts of the tiles before it by
// decoupled look-back (tiles take their index from a ticket, so every predecessor is already running), reorders the
// tile in shared memory and writes each digit's run to its final place with coalesced stores.  Pass 0 has no payload
// to read: the payload is the element index.  A pass at or above the key width returns at once.
template <int ITEMS>
__global__ void __launch_bounds__(SORT_THREADS, (ITEMS > 8 ? 3 : 4)) onesweep_kernel(SortView sv, PlanView P, const uint32_t *__restrict__ hist,
                                                                uint32_t *__restrict__ state, uint32_t *__restrict__ ticket,
                                                                int pass) {
    chain_sync();
    const int shift = pass * RADIX_BITS;
    if ((uint32_t)shift >= __ldg(&sv.scalars[0])) return;
    const bool odd = (pass & 1) != 0;
    const uint32_t *__restrict__ kin = odd ? sv.k[1] : sv.k[0];
    const uint32_t *__restrict__ vin = odd ? sv.v[1] : sv.v[0];
    uint32_t *__restrict__ kout = const_cast<uint32_t *>(odd ? sv.k[0] : sv.k[1]);
    uint32_t *__restrict__ vout = const_cast<uint32_t *>(odd ? sv.v[0] : sv.v[1]);
    constexpr int NW = SORT_THREADS / 32;
    constexpr int TILE = SORT_THREADS * ITEMS;
    __shared__ uint32_t wcnt[NW][RADIX];
    __shared__ uint32_t gbase[RADIX], lstart[RADIX];
    __shared__ uint32_t skey[TILE], sval[TILE];
    __shared__ uint32_t wsum[NW + 1];
    __shared__ uint32_t s_tile;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < NW * RADIX; i += SORT_THREADS) (&wcnt[0][0])[i] = 0u;
    __syncthreads();
    const uint32_t tile = s_tile;
    const TileDesc t = get_tile(P, tile);
    const SegDesc sg = get_seg(P, t.seg);
    // all loads of the tile are in flight before the first rank is computed
    uint32_t key[ITEMS], val[ITEMS];
    uint16_t loc[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const uint32_t e = w * (32 * ITEMS) + r * 32 + l;
        key[r] = (e < t.count) ? __ldcs(&kin[t.begin + e]) : 0xFFFFFFFFu;
    }
    if (pass == 0) {
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) val[r] = t.begin + w * (32 * ITEMS) + r * 32 + l;
    } else {
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const uint32_t e = w * (32 * ITEMS) + r * 32 + l;
            val[r] = (e < t.count) ? __ldcs(&vin[t.begin + e]) : 0u;
        }
    }
    const uint32_t lt = (1u << l) - 1u;
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const uint32_t e = w * (32 * ITEMS) + r * 32 + l;
        const bool valid = e < t.count;
        const uint32_t digit = valid ? ((key[r] >> shift) & (RADIX - 1)) : (uint32_t)RADIX;
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (valid && l == leader) {
            base = wcnt[w][digit];
            wcnt[w][digit] = base + __popc(peers);
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        loc[r] = (uint16_t)(base + __popc(peers & lt));
        __syncwarp();
    }
    __syncthreads();
    // thread d owns digit d from here on
    const uint32_t d = threadIdx.x;
    uint32_t cnt_d = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) { const uint32_t c = wcnt[i][d]; wcnt[i][d] = cnt_d; cnt_d += c; }
    uint32_t *st = state + (size_t)tile * RADIX;
    uint32_t excl = 0;
    if (t.tile_in_seg == 0) {
        st_vol_g(&st[d], cnt_d | LB_INCL);
    } else {
        st_vol_g(&st[d], cnt_d | LB_PART);
        // Look-back, LB words at a time: the loads of a batch are independent (all in flight together), so a tile that
        // starts together with its predecessors (small inputs: every tile is resident at once) pays one L2 round trip
        // per LB predecessors instead of one per predecessor.  The cloud's first tile always publishes an inclusive
        // count, so the walk never leaves the cloud; words below it read as "inclusive 0".
        constexpr int LB = 16;
        const int jmin = (int)sg.tile_begin;
        int j = (int)tile - 1;
        while (true) {
            uint32_t v[LB];
#pragma unroll
            for (int q = 0; q < LB; ++q) v[q] = (j - q >= jmin) ? ld_vol_g(&state[(size_t)(j - q) * RADIX + d]) : LB_INCL;
            bool done = false;
            int used = 0;
#pragma unroll
            for (int q = 0; q < LB; ++q) {
                if (!done && used == q) {
                    if ((v[q] & ~LB_MASK) != 0u) {
                        excl += v[q] & LB_MASK;
                        ++used;
                        if (v[q] & LB_INCL) done = true;
                    }
                }
            }
            if (done) break;
            j -= used;                      // re-poll from the first word that was not ready
        }
        st_vol_g(&st[d], (excl + cnt_d) | LB_INCL);
    }
    // digit bases: cloud-wide (histogram) and tile-local
    const uint32_t hd = __ldg(&hist[((size_t)t.seg * 4 + pass) * RADIX + d]);
    const uint32_t hex = block_excl_scan(hd, wsum, nullptr);
    const uint32_t lex = block_excl_scan(cnt_d, wsum, nullptr);
    gbase[d] = sg.begin + hex + excl;
    lstart[d] = lex;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const uint32_t e = w * (32 * ITEMS) + r * 32 + l;
        if (e < t.count) {
            const uint32_t digit = (key[r] >> shift) & (RADIX - 1);
            const uint32_t pos = lstart[digit] + wcnt[w][digit] + loc[r];
            skey[pos] = key[r];
            sval[pos] = val[r];
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t pos = i * SORT_THREADS + threadIdx.x;
        if (pos < t.count) {
            const uint32_t k = skey[pos];
            const uint32_t digit = (k >> shift) & (RADIX - 1);
            const uint32_t g = gbase[digit] + (pos - lstart[digit]);
            kout[g] = k;
            vout[g] = sval[pos];
        }
    }
}

// ---- kernel 4: run heads (one run = one occupied voxel) in one kernel: per-tile head flags (blocked arrangement,
// thread t owns ITEMS consecutive keys), a decoupled look-back over the tiles' head counts (one chain over the whole
// batch: run indices are global), then the compacted run table {first element, cloud}, the per-cloud first-run table
// and the total.
template <int ITEMS>
__global__ void __launch_bounds__(SORT_THREADS) runs_kernel(SortView sv, PlanView P, const VoxLayout *__restrict__ layouts,
                                                            uint32_t *__restrict__ state, uint32_t *__restrict__ ticket,
                                                            uint32_t *__restrict__ run_start, uint32_t *__restrict__ run_seg,
                                                            uint32_t *__restrict__ run_seg_off, uint32_t *__restrict__ scalars) {
    chain_sync();
    const uint32_t *__restrict__ keys = sv.keys();
    __shared__ uint32_t s_tile, s_ok, s_inv, s_excl;
    __shared__ uint32_t wsum[SORT_THREADS / 32 + 1];
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const TileDesc t = get_tile(P, tile);
    const SegDesc sg = get_seg(P, t.seg);
    if (threadIdx.x == 0) { s_ok = layouts[t.seg].ok; s_inv = layouts[t.seg].ncells; }
    __syncthreads();
    const uint32_t e0 = threadIdx.x * ITEMS;
    uint32_t flags = 0;
    if (s_ok && e0 < t.count) {
        const uint32_t gi = t.begin + e0;
        uint32_t prev = (gi > sg.begin) ? keys[gi - 1] : 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            if (e0 + j < t.count) {
                const uint32_t k = keys[gi + j];
                const bool head = (k != s_inv) && ((gi + j == sg.begin) || (k != prev));
                flags |= head ? (1u << j) : 0u;
                prev = k;
            }
        }
    }
    const uint32_t cnt = __popc(flags);
    uint32_t total;
    const uint32_t lex = block_excl_scan(cnt, wsum, &total);
    if (threadIdx.x < 32) {
        // warp-parallel look-back: lane i inspects tile - 1 - i (a virtual tile before tile 0 holds an inclusive 0)
        const int l = threadIdx.x;
        uint32_t excl = 0;
        if (tile == 0) {
            if (l == 0) st_vol_g(&state[0], total | LB_INCL);
        } else {
            if (l == 0) st_vol_g(&state[tile], total | LB_PART);
            int base = (int)tile;
            while (true) {
                const int j = base - 1 - l;
                uint32_t v = LB_INCL;
                if (j >= 0) { do { v = ld_vol_g(&state[j]); } while ((v & ~LB_MASK) == 0u); }
                const uint32_t incl_mask = __ballot_sync(0xffffffffu, (v & LB_INCL) != 0u);
                const int stop = __ffs(incl_mask) - 1;                    // nearest predecessor with an inclusive count
                uint32_t c = (stop < 0 || l <= stop) ? (v & LB_MASK) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                excl += c;
                if (stop >= 0) break;
                base -= 32;
            }
            if (l == 0) st_vol_g(&state[tile], (excl + total) | LB_INCL);
        }
        if (l == 0) s_excl = excl;
    }
    __syncthreads();
    const uint32_t excl = s_excl;
    uint32_t o = excl + lex;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
        if (flags & (1u << j)) { run_start[o] = t.begin + e0 + j; run_seg[o] = t.seg; ++o; }
    if (threadIdx.x == 0) {
        if (t.tile_in_seg == 0) {
            // first tile of a cloud: the first-run entry of this cloud and of the empty clouds just before it
            uint32_t s = t.seg;
            run_seg_off[s] = excl;
            while (s > 0 && get_seg(P, s - 1).count == 0u) { --s; run_seg_off[s] = excl; }
        }
        if (tile == P.ntiles - 1u) {
            for (uint32_t s = t.seg + 1; s <= P.B; ++s) run_seg_off[s] = excl + total;
            scalars[1] = excl + total;
        }
    }
}

// no tiles at all (every cloud of the batch is empty): layouts, totals and the per-cloud run table
__global__ void empty_plan_kernel(VoxLayout *layouts, uint32_t *scalars, uint32_t *run_seg_off, uint32_t B, float lx, float ly,
                                  float lz) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < B) {
        const float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        VoxLayout L;
        compute_layout(mn, mx, 0u, lx, ly, lz, L);
        layouts[s] = L;
    }
    if (s <= B) run_seg_off[s] = 0u;
    if (s == 0) { scalars[0] = 0u; scalars[1] = 0u; scalars[10] = 0u; scalars[11] = 0u; }
}

__global__ void prepared_layout_kernel(VoxLayout *layouts, uint32_t *scalars, uint32_t *hist, uint32_t hist_words, uint32_t B,
                                       uint32_t invalid_key, int nbits) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t i = s; i < B; i += gridDim.x * blockDim.x) {
        VoxLayout L;
        memset(&L, 0, sizeof(L));
        L.ok = 1; L.ncells = invalid_key; L.nbits = nbits;
        layouts[i] = L;
    }
    for (uint32_t i = s; i < hist_words; i += gridDim.x * blockDim.x) hist[i] = 0u;
    if (s == 0) {
        scalars[0] = (uint32_t)nbits; scalars[1] = 0u; scalars[2] = 0u;
        for (int i = 4; i < 12; ++i) scalars[i] = 0u;
    }
}

// ------------------------------------------------------------------ VoxPipeline -------------
int VoxPipeline::plan(const uint32_t *off, size_t nB, cudaStream_t st) {
    if ((size_t)off[nB] >= B2_MAX_POINTS) { set_error("cloud too large: %u points (limit %zu per call)", off[nB], (size_t)B2_MAX_POINTS); return B2_ERR_INVALID; }
    const bool same = (nB == B && nB > 1 && h_off.size() == nB + 1 && memcmp(h_off.data(), off, (nB + 1) * 4) == 0);
    B = nB;
    N = off[nB];
    // small tiles only while they are needed to fill the SMs: the look-back of a radix pass walks over the tiles that
    // started together, so fewer, larger tiles are cheaper as soon as there are enough of them
    static const size_t large_from = [] { const char *e = getenv("B2_SORT_LARGE_FROM"); return e ? (size_t)atoll(e) : SORT_LARGE_FROM; }();
    tile_elems = SORT_THREADS * (N >= large_from ? SORT_ITEMS_LARGE : SORT_ITEMS_SMALL);
    int rc;
    if (nB <= 1) {
        ntiles = (N + tile_elems - 1) / tile_elems;
        h_off.clear();
    } else if (!same) {
        h_tiles.clear();
        h_segs.resize(nB);
        for (size_t s = 0; s < nB; ++s) {
            uint32_t b = off[s], c = off[s + 1] - off[s];
            SegDesc sd;
            sd.begin = b; sd.count = c; sd.tile_begin = (uint32_t)h_tiles.size();
            sd.ntiles = (c + tile_elems - 1) / tile_elems;
            for (uint32_t t = 0; t < sd.ntiles; ++t) {
                TileDesc td;
                td.seg = (uint32_t)s; td.begin = b + t * tile_elems;
                td.count = (c - t * tile_elems < tile_elems) ? c - t * tile_elems : tile_elems;
                td.tile_in_seg = t;
                h_tiles.push_back(td);
            }
            h_segs[s] = sd;
        }
        ntiles = h_tiles.size();
        if ((rc = d_tiles.reserve((ntiles + 1) * sizeof(TileDesc)))) return rc;
        if ((rc = d_segs.reserve((B + 1) * sizeof(SegDesc)))) return rc;
        if (ntiles) B2_CUDA(cudaMemcpyAsync(d_tiles.p, h_tiles.data(), ntiles * sizeof(TileDesc), cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemcpyAsync(d_segs.p, h_segs.data(), B * sizeof(SegDesc), cudaMemcpyHostToDevice, st));
        // the async copies read h_tiles / h_segs (pageable): the runtime stages them before returning
        h_off.assign(off, off + nB + 1);
    }
    if ((rc = d_layouts.reserve((B + 1) * sizeof(VoxLayout)))) return rc;
    if (d_scalars.cap < 16 * sizeof(uint32_t)) scalars_ready = false;
    if ((rc = d_scalars.reserve(16 * sizeof(uint32_t)))) return rc;
    if (!scalars_ready) {
        // the tickets are self-resetting from here on (bbox_layout_kernel / prepared_layout_kernel)
        B2_CUDA(cudaMemsetAsync(d_scalars.p, 0, 16 * sizeof(uint32_t), st));
        scalars_ready = true;
    }
    if ((rc = d_hist.reserve((B + 1) * 4 * RADIX * sizeof(uint32_t)))) return rc;
    for (int i = 0; i < 2; ++i) {
        if ((rc = d_keys[i].reserve((N + 1) * sizeof(uint32_t)))) return rc;
        if ((rc = d_vals[i].reserve((N + 1) * sizeof(uint32_t)))) return rc;
    }
    if ((rc = d_bbox_part.reserve((ntiles + 1) * 8 * sizeof(float)))) return rc;
    if ((rc = d_state.reserve(((size_t)ntiles * (4 * RADIX + 1) + 16) * sizeof(uint32_t)))) return rc;
    if ((rc = d_run_start.reserve((N + 2) * sizeof(uint32_t)))) return rc;
    if ((rc = d_run_seg.reserve((N + 2) * sizeof(uint32_t)))) return rc;
    if ((rc = d_run_seg_off.reserve((B + 2) * sizeof(uint32_t)))) return rc;
    return 0;
}

int VoxPipeline::sort_and_runs(int npass_launch, cudaStream_t st) {
    const uint32_t nt = (uint32_t)ntiles;
    const SortView sv = view();
    const PlanView P = plan_view();
    uint32_t *sc = d_scalars.as<uint32_t>();
    uint32_t *state = d_state.as<uint32_t>();
    const bool large = tile_elems == SORT_THREADS * SORT_ITEMS_LARGE;
    for (int p = 0; p < npass_launch; ++p) {
        uint32_t *stp = state + (size_t)p * nt * RADIX;
        if (large) launch_chain(onesweep_kernel<SORT_ITEMS_LARGE>, nt, SORT_THREADS, 0, st, sv, P, d_hist.as<uint32_t>(), stp, sc + 4 + p, p);
        else launch_chain(onesweep_kernel<SORT_ITEMS_SMALL>, nt, SORT_THREADS, 0, st, sv, P, d_hist.as<uint32_t>(), stp, sc + 4 + p, p);
        B2_LAUNCH_CHECK();
    }
    uint32_t *str = state + (size_t)4 * nt * RADIX;
    if (large)
        launch_chain(runs_kernel<SORT_ITEMS_LARGE>, nt, SORT_THREADS, 0, st, sv, P, layouts(), str, sc + 8, d_run_start.as<uint32_t>(),
                     d_run_seg.as<uint32_t>(), d_run_seg_off.as<uint32_t>(), sc);
    else
        launch_chain(runs_kernel<SORT_ITEMS_SMALL>, nt, SORT_THREADS, 0, st, sv, P, layouts(), str, sc + 8, d_run_start.as<uint32_t>(),
                     d_run_seg.as<uint32_t>(), d_run_seg_off.as<uint32_t>(), sc);
    B2_LAUNCH_CHECK();
    return 0;
}

template <class Op>
static int run_pipeline(VoxPipeline &pp, const float4 *d_pts, float4 *d_pts_out, const Op &op, float lx, float ly, float lz,
                        int nbits_hint, cudaStream_t st, int (VoxPipeline::*tail)(int, cudaStream_t)) {
    const uint32_t nB = (uint32_t)pp.B, nt = (uint32_t)pp.ntiles;
    uint32_t *sc = pp.d_scalars.as<uint32_t>();
    VoxLayout *lay = pp.d_layouts.as<VoxLayout>();
    if (nt == 0) {
        empty_plan_kernel<<<(nB + 1 + 127) / 128, 128, 0, st>>>(lay, sc, pp.d_run_seg_off.as<uint32_t>(), nB, lx, ly, lz);
        B2_LAUNCH_CHECK();
        return 0;
    }
    const PlanView P = pp.plan_view();
    launch_chain(bbox_layout_kernel<Op>, nt, SORT_THREADS, 0, st, d_pts, d_pts_out, op, P, lx, ly, lz, pp.d_bbox_part.as<float>(), lay, sc,
                 pp.d_hist.as<uint32_t>(), nB * 4 * RADIX);
    B2_LAUNCH_CHECK();
    launch_chain(key_hist_kernel<true>, nt, SORT_THREADS, 0, st, (const float4 *)(d_pts_out ? d_pts_out : d_pts), P, lay, pp.d_keys[0].as<uint32_t>(),
                 pp.d_hist.as<uint32_t>(), pp.d_state.as<uint32_t>(), sc);
    B2_LAUNCH_CHECK();
    int passes = 32 / RADIX_BITS;
    if (nbits_hint > 0) passes = (nbits_hint + RADIX_BITS - 1) / RADIX_BITS;
    return (pp.*tail)(passes, st);
}

int VoxPipeline::run(const float4 *d_pts, float lx, float ly, float lz, int nbits_hint, cudaStream_t st) {
    return run_pipeline(*this, d_pts, nullptr, NoIngest(), lx, ly, lz, nbits_hint, st, &VoxPipeline::sort_and_runs);
}

int VoxPipeline::run_ingest(const float4 *d_raw, float4 *d_ingested, const IngestOp &op, float lx, float ly, float lz, cudaStream_t st) {
    return run_pipeline(*this, d_raw, d_ingested, op, lx, ly, lz, 0, st, &VoxPipeline::sort_and_runs);
}

int VoxPipeline::run_prepared(uint32_t invalid_key, int nbits, cudaStream_t st) {
    const uint32_t nB = (uint32_t)B, nt = (uint32_t)ntiles;
    uint32_t *sc = d_scalars.as<uint32_t>();
    VoxLayout *lay = d_layouts.as<VoxLayout>();
    if (nt == 0) {
        empty_plan_kernel<<<(nB + 1 + 127) / 128, 128, 0, st>>>(lay, sc, d_run_seg_off.as<uint32_t>(), nB, 1.f, 1.f, 1.f);
        B2_LAUNCH_CHECK();
        return 0;
    }
    const PlanView P = plan_view();
    prepared_layout_kernel<<<(nB * 4 * RADIX + 255) / 256 < 64 ? (nB * 4 * RADIX + 255) / 256 : 64, 256, 0, st>>>(
        lay, sc, d_hist.as<uint32_t>(), nB * 4 * RADIX, nB, invalid_key, nbits);
    B2_LAUNCH_CHECK();
    launch_chain(key_hist_kernel<false>, nt, SORT_THREADS, 0, st, (const float4 *)nullptr, P, lay, d_keys[0].as<uint32_t>(), d_hist.as<uint32_t>(),
                 d_state.as<uint32_t>(), sc);
    B2_LAUNCH_CHECK();
    return sort_and_runs((nbits + RADIX_BITS - 1) / RADIX_BITS, st);
}

void VoxPipeline::release() {
    d_tiles.release(); d_segs.release(); d_layouts.release(); d_scalars.release(); d_hist.release();
    for (int i = 0; i < 2; ++i) { d_keys[i].release(); d_vals[i].release(); }
    d_bbox_part.release(); d_state.release();
    d_run_start.release(); d_run_seg.release(); d_run_seg_off.release();
    scalars_ready = false;
}

// ------------------------------------------------------------------ voxel filter ------------
// Centroid of every occupied voxel = the float sum of its member points in input order (the stable sort keeps it),
// formed left to right exactly as pcl::VoxelGrid does for that member order, then divided by the count.
//   vf_centroid_kernel: a warp takes 32 consecutive voxels; a voxel of up to VF_SEQ points is folded by ONE lane
//     (32 voxels in parallel, gathers four deep); crowded voxels are only appended to a work list;
//   vf_crowded_kernel: one warp per crowded voxel: the lanes gather 128 members at a time, two groups ahead of the
//     fold (index load -> point load -> fold are software-pipelined), every lane folds them through shuffles, so the
//     cost is the order-exact add chain itself (4 cycles per member).
// Output in ascending voxel index.
constexpr uint32_t VF_SEQ = 32;
__device__ __forceinline__ void run_bounds(uint32_t j, const uint32_t *__restrict__ run_start, const uint32_t *__restrict__ run_seg,
                                           const uint32_t *__restrict__ run_seg_off, const PlanView &P,
                                           const VoxLayout *__restrict__ layouts, uint32_t &s, uint32_t &e) {
    const uint32_t sg = run_seg[j];
    s = run_start[j];
    e = (j + 1 < run_seg_off[sg + 1]) ? run_start[j + 1] : get_seg(P, sg).begin + layouts[sg].n_finite;
}

__global__ void __launch_bounds__(256) vf_centroid_kernel(const float4 *__restrict__ pts, SortView sv,
                                                          const uint32_t *__restrict__ run_start,
                                                          const uint32_t *__restrict__ run_seg,
                                                          const uint32_t *__restrict__ run_seg_off, PlanView P,
                                                          const VoxLayout *__restrict__ layouts, uint32_t *__restrict__ crowded,
                                                          uint32_t *__restrict__ n_crowded, const uint32_t *__restrict__ out_base,
                                                          uint32_t out_capacity,
                                                          float4 *__restrict__ out, int32_t *__restrict__ out_idx,
                                                          int32_t *__restrict__ out_cnt) {
    chain_sync();
    const uint32_t total = sv.scalars[1];
    // append mode (batched ingest): the output starts at the cursor another call left on the device; a batch that
    // would not fit is dropped as a whole (vf_append_kernel raises the overflow flag)
    const uint32_t ob = out_base ? *out_base : 0u;
    if ((size_t)ob + total > (size_t)out_capacity) return;
    out += ob;
    const uint32_t *__restrict__ keys = sv.keys();
    const uint32_t *__restrict__ vals = sv.vals();
    const int l = threadIdx.x & 31;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    // consecutive groups of 32 voxels go to different CTAs (a small cloud still spreads over the SMs)
    for (uint32_t j0 = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32u; j0 < total; j0 += nwarps * 32u) {
        const uint32_t j = j0 + l;
        const bool valid = j < total;
        uint32_t s = 0, e = 0;
        if (valid) run_bounds(j, run_start, run_seg, run_seg_off, P, layouts, s, e);
        const uint32_t len = e - s;
        const bool is_crowded = len > VF_SEQ;
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
            float ax = 0.f, ay = 0.f, az = 0.f, ai = 0.f;
            uint32_t k = s;
            for (; k + 4 <= e; k += 4) {
                const uint32_t v0 = vals[k], v1 = vals[k + 1], v2 = vals[k + 2], v3 = vals[k + 3];
                const float4 p0 = __ldg(&pts[v0]), p1 = __ldg(&pts[v1]), p2 = __ldg(&pts[v2]), p3 = __ldg(&pts[v3]);
                ax = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(ax, p0.x), p1.x), p2.x), p3.x);
                ay = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(ay, p0.y), p1.y), p2.y), p3.y);
                az = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(az, p0.z), p1.z), p2.z), p3.z);
                ai = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(ai, p0.w), p1.w), p2.w), p3.w);
            }
            for (; k < e; ++k) {
                const float4 p = __ldg(&pts[vals[k]]);
                ax = __fadd_rn(ax, p.x); ay = __fadd_rn(ay, p.y); az = __fadd_rn(az, p.z); ai = __fadd_rn(ai, p.w);
            }
            const float n = (float)len;
            out[j] = make_float4(__fdiv_rn(ax, n), __fdiv_rn(ay, n), __fdiv_rn(az, n), __fdiv_rn(ai, n));
            if (out_idx) out_idx[j] = (int32_t)keys[s];
            if (out_cnt) out_cnt[j] = (int32_t)len;
        }
    }
}

constexpr int CROWD_THREADS = 128;     // 4 warps, each with a ring of CROWD_RING staging groups of CROWD_GROUP points
constexpr int CROWD_GROUP = 128;
constexpr int CROWD_RING = 4;
__global__ void __launch_bounds__(CROWD_THREADS) vf_crowded_kernel(const float4 *__restrict__ pts, SortView sv,
                                                                   const uint32_t *__restrict__ run_start,
                                                                   const uint32_t *__restrict__ run_seg,
                                                                   const uint32_t *__restrict__ run_seg_off, PlanView P,
                                                                   const VoxLayout *__restrict__ layouts,
                                                                   const uint32_t *__restrict__ crowded,
                                                                   const uint32_t *__restrict__ n_crowded,
                                                                   const uint32_t *__restrict__ out_base, uint32_t out_capacity,
                                                                   float4 *__restrict__ out, int32_t *__restrict__ out_idx,
                                                                   int32_t *__restrict__ out_cnt) {
    chain_sync();
    const uint32_t n = *n_crowded;
    const uint32_t ob = out_base ? *out_base : 0u;
    if ((size_t)ob + sv.scalars[1] > (size_t)out_capacity) return;
    out += ob;
    const uint32_t *__restrict__ keys = sv.keys();
    const uint32_t *__restrict__ vals = sv.vals();
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    __shared__ float4 stage[CROWD_THREADS / 32][CROWD_RING][CROWD_GROUP];
    for (uint32_t i = w * gridDim.x + blockIdx.x; i < n; i += nwarps) {
        const uint32_t j = crowded[i];
        uint32_t s, e;
        run_bounds(j, run_start, run_seg, run_seg_off, P, layouts, s, e);
        // indices of a group of 128 members (four per lane)
        auto ldv = [&](uint32_t c, uint32_t (&v)[4]) {
#pragma unroll
            for (int d = 0; d < 4; ++d) { const uint32_t idx = c + d * 32 + l; v[d] = (idx < e) ? __ldg(&vals[idx]) : 0xFFFFFFFFu; }
        };
        // the group's points: asynchronous 16-byte copies straight into ring slot `buf` (one commit group per call)
        auto gather = [&](const uint32_t (&v)[4], int buf) {
#pragma unroll
            for (int d = 0; d < 4; ++d)
                if (v[d] != 0xFFFFFFFFu) cp_async16(&stage[w][buf][d * 32 + l], &pts[v[d]]);
            cp_async_commit();
        };
        // lane c (0..3) folds component c: one 4-byte shared load + one FADD per member, i.e. the order-exact add
        // chain itself (4 cycles per member) is the cost
        float acc = 0.f;
        const int comp = l & 3;
        auto fold = [&](int buf, uint32_t c) {
            const float *q = reinterpret_cast<const float *>(stage[w][buf]) + comp;
            const int m = (e - c < (uint32_t)CROWD_GROUP) ? (int)(e - c) : CROWD_GROUP;
            if (m == CROWD_GROUP) {
#pragma unroll 32
                for (int k = 0; k < CROWD_GROUP; ++k) acc = __fadd_rn(acc, q[4 * k]);
            } else {
                // most crowded voxels have fewer members than a group: sixteen at a time, then the rest
                int k = 0;
                for (; k + 16 <= m; k += 16) {
#pragma unroll
                    for (int u = 0; u < 16; ++u) acc = __fadd_rn(acc, q[4 * (k + u)]);
                }
#pragma unroll 4
                for (; k < m; ++k) acc = __fadd_rn(acc, q[4 * k]);
            }
        };
        // Software pipeline over groups: while group g is folded, the points of g+1 .. g+3 are in flight into the
        // ring and the indices of g+4 into registers.
        uint32_t v0[4], v1[4], v2[4], vn[4];
        ldv(s, v0); ldv(s + CROWD_GROUP, v1); ldv(s + 2 * CROWD_GROUP, v2); ldv(s + 3 * CROWD_GROUP, vn);
        gather(v0, 0); gather(v1, 1); gather(v2, 2);
        int buf = 0;
        for (uint32_t c = s; c < e; c += CROWD_GROUP) {
            gather(vn, (buf + 3) & (CROWD_RING - 1));            // group g+3 (an empty commit group past the end)
            ldv(c + 4 * CROWD_GROUP, vn);
            cp_async_wait<3>();                                   // group g has landed (this lane's copies) ...
            __syncwarp();                                         // ... and every other lane's
            fold(buf, c);
            __syncwarp();                                         // slot `buf` is rewritten three groups from now
            buf = (buf + 1) & (CROWD_RING - 1);
        }
        cp_async_wait<0>();
        const float ax = __shfl_sync(0xffffffffu, acc, 0), ay = __shfl_sync(0xffffffffu, acc, 1);
        const float az = __shfl_sync(0xffffffffu, acc, 2), ai = __shfl_sync(0xffffffffu, acc, 3);
        if (l == 0) {
            const float nn = (float)(e - s);
            out[j] = make_float4(__fdiv_rn(ax, nn), __fdiv_rn(ay, nn), __fdiv_rn(az, nn), __fdiv_rn(ai, nn));
            if (out_idx) out_idx[j] = (int32_t)keys[s];
            if (out_cnt) out_cnt[j] = (int32_t)(e - s);
        }
    }
}

// append mode: publish the per-cloud offsets of this batch relative to the whole output and advance the cursor
// (cursor[0] = points in the output so far, cursor[1] = overflow flag)
__global__ void __launch_bounds__(256) vf_append_kernel(const uint32_t *__restrict__ run_seg_off, uint32_t B,
                                                        uint32_t *__restrict__ out_offsets, uint32_t *__restrict__ cursor,
                                                        uint32_t out_capacity) {
    __shared__ uint32_t s_cur, s_tot;
    if (threadIdx.x == 0) { s_cur = cursor[0]; s_tot = run_seg_off[B]; }
    __syncthreads();
    const uint32_t cur = s_cur;
    const bool fits = (size_t)cur + s_tot <= (size_t)out_capacity;
    for (uint32_t s = threadIdx.x; s <= B; s += blockDim.x) out_offsets[s] = cur + (fits ? run_seg_off[s] : 0u);
    if (threadIdx.x == 0) { if (fits) cursor[0] = cur + s_tot; else cursor[1] = 1u; }
}

}  // namespace b2

// ------------------------------------------------------------------ C ABI: b2vf_* ------------
using namespace b2;

struct b2vf {
    int device = 0;
    float leaf[3];
    cudaStream_t own = nullptr, st = nullptr;
    VoxPipeline pipe;
    DevBuf d_in, d_out, d_idx, d_cnt, d_crowded;
    PinBuf h_in, h_out, h_idx, h_cnt, h_misc;
};

extern "C" const char *b2_last_error(void) { return b2::last_error(); }
extern "C" uint64_t b2_kernel_launch_count(void) { return b2::g_launches.load(); }
extern "C" int b2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

namespace b2 {
int check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); libb2ndt has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count=0");
        return B2_ERR_CUDA;
    }
    if (device < 0 || device >= n) { set_error("device %d out of range (have %d)", device, n); return B2_ERR_INVALID; }
    cudaDeviceProp prop;
    B2_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libb2ndt is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return B2_ERR_CUDA;
    }
    return 0;
}
}  // namespace b2

extern "C" int b2vf_create(float lx, float ly, float lz, int device, b2vf **out) {
    if (!out) { set_error("b2vf_create: out is NULL"); return B2_ERR_INVALID; }
    *out = nullptr;
    if (!(lx > 0.f) || !(ly > 0.f) || !(lz > 0.f)) { set_error("b2vf_create: leaf sizes must be > 0"); return B2_ERR_INVALID; }
    int rc = check_device(device);
    if (rc) return rc;
    B2_CUDA(cudaSetDevice(device));
    b2vf *h = new b2vf();
    h->device = device;
    h->leaf[0] = lx; h->leaf[1] = ly; h->leaf[2] = lz;
    cudaError_t e = cudaStreamCreateWithFlags(&h->own, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e)); delete h; return B2_ERR_CUDA; }
    h->st = h->own;
    *out = h;
    return 0;
}

extern "C" void b2vf_destroy(b2vf *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->st);
    h->pipe.release();
    h->d_in.release(); h->d_out.release(); h->d_idx.release(); h->d_cnt.release(); h->d_crowded.release();
    h->h_in.release(); h->h_out.release(); h->h_idx.release(); h->h_cnt.release(); h->h_misc.release();
    if (h->own) cudaStreamDestroy(h->own);
    delete h;
}

extern "C" int b2vf_set_stream(b2vf *h, void *stream) {
    if (!h) { set_error("b2vf_set_stream: NULL handle"); return B2_ERR_INVALID; }
    h->st = stream ? (cudaStream_t)stream : h->own;
    return 0;
}

static int vf_launch_centroids(b2vf *h, const float4 *d_in, float4 *d_out, int32_t *d_idx, int32_t *d_cnt, size_t n,
                               const uint32_t *d_out_base = nullptr, uint32_t out_capacity = 0xFFFFFFFFu) {
    if (n == 0) return 0;
    int rc;
    if ((rc = h->d_crowded.reserve((n / VF_SEQ + 64) * 4))) return rc;
    // a warp takes 32 voxels at a time; enough warps for a typical cloud, grid-stride beyond that; multiple of 148 SMs
    unsigned blocks = (unsigned)((n / 256 + 147) / 148) * 148u;
    if (blocks < 148) blocks = 148;
    if (blocks > 148 * 16) blocks = 148 * 16;
    uint32_t *ncr = const_cast<uint32_t *>(h->pipe.scalars()) + 10;
    launch_chain(vf_centroid_kernel, blocks, 256, 0, h->st, d_in, h->pipe.view(), h->pipe.run_start(), h->pipe.run_seg(), h->pipe.run_seg_off(),
                 h->pipe.plan_view(), h->pipe.layouts(), h->d_crowded.as<uint32_t>(), ncr, d_out_base,
                 out_capacity, d_out, d_idx, d_cnt);
    B2_LAUNCH_CHECK();
    launch_chain(vf_crowded_kernel, 148 * 8, CROWD_THREADS, 0, h->st, d_in, h->pipe.view(), h->pipe.run_start(), h->pipe.run_seg(), h->pipe.run_seg_off(),
                 h->pipe.plan_view(), h->pipe.layouts(), h->d_crowded.as<uint32_t>(), ncr, d_out_base,
                 out_capacity, d_out, d_idx, d_cnt);
    B2_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2vf_filter(b2vf *h, const void *in, size_t n, size_t stride, size_t ioff, void *out, size_t out_capacity,
                           size_t out_stride, size_t out_ioff, size_t *m, int32_t *out_idx, int32_t *out_cnt) {
    if (!h || !m) { set_error("b2vf_filter: NULL handle or m"); return B2_ERR_INVALID; }
    *m = 0;
    if (n == 0) return 0;
    if (!in || !out) { set_error("b2vf_filter: NULL cloud pointer"); return B2_ERR_INVALID; }
    if (stride < 16 || ioff + 4 > stride || out_stride < 16 || out_ioff + 4 > out_stride || ((stride | ioff | out_stride | out_ioff) & 3)) {
        set_error("b2vf_filter: bad stride / intensity offset"); return B2_ERR_INVALID;
    }
    if (n >= B2_MAX_POINTS) { set_error("b2vf_filter: cloud too large"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(h->device));
    int rc;
    if ((rc = h->h_in.reserve(n * 16))) return rc;
    if ((rc = h->d_in.reserve(n * 16))) return rc;
    if ((rc = h->d_out.reserve(n * 16))) return rc;
    if ((rc = h->d_idx.reserve(n * 4))) return rc;
    if ((rc = h->d_cnt.reserve(n * 4))) return rc;
    if ((rc = h->h_misc.reserve(256))) return rc;
    // host pass: repack to float4 and (for free) bound the key width so the sort runs only the radix
    // passes it needs
    int nbits_hint = 0;
    const void *h2d_from = h->h_in.p;
    if (stride == 16 && ioff == 12 && n > 1000000) {
        // large packed cloud: DMA straight from the caller's memory (full PCIe rate when it is pinned, the
        // driver's staging otherwise); the key width is then decided on the device
        h2d_from = in;
    } else if (n > 400000) {
        // big cloud: the (threaded) repack also bounds the key width, which saves whole radix passes over N keys
        float mn[3], mx[3];
        pack_cloud_f4_bbox(in, n, stride, ioff, h->h_in.as<float>(), mn, mx);
        nbits_hint = key_bits_from_bbox(mn, mx, h->leaf[0], h->leaf[1], h->leaf[2]);
    } else {
        // a scan: a plain (vectorisable) repack is cheaper than a bounding-box pass on one host core, and a radix
        // pass over ~1e5 keys costs microseconds; the key width is decided on the device
        pack_cloud_f4(in, n, stride, ioff, h->h_in.as<float>());
    }
    B2_CUDA(cudaMemcpyAsync(h->d_in.p, h2d_from, n * 16, cudaMemcpyHostToDevice, h->st));
    uint32_t off[2] = {0u, (uint32_t)n};
    if ((rc = h->pipe.plan(off, 1, h->st))) return rc;
    if ((rc = h->pipe.run(h->d_in.as<float4>(), h->leaf[0], h->leaf[1], h->leaf[2], nbits_hint, h->st))) return rc;
    if ((rc = vf_launch_centroids(h, h->d_in.as<float4>(), h->d_out.as<float4>(), h->d_idx.as<int32_t>(),
                                  h->d_cnt.as<int32_t>(), n))) return rc;
    uint32_t *misc = h->h_misc.as<uint32_t>();
    B2_CUDA(cudaMemcpyAsync(misc, h->pipe.scalars(), 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaMemcpyAsync(misc + 8, h->pipe.layouts(), sizeof(VoxLayout), cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaStreamSynchronize(h->st));
    VoxLayout L;
    memcpy(&L, misc + 8, sizeof(L));
    size_t M = misc[1];
    char *o = (char *)out;
    auto put = [&](size_t j, const float *p4) {
        float *q = (float *)(o + j * out_stride);
        if (out_stride >= 32) memset(q, 0, out_stride);
        q[0] = p4[0]; q[1] = p4[1]; q[2] = p4[2];
        if (out_stride >= 32 || out_ioff != 12) q[3] = 1.0f;   // PointXYZI padding data[3]
        *(float *)(o + j * out_stride + out_ioff) = p4[3];
    };
    if (!L.ok) {
        if (L.n_finite == 0) { *m = 0; return 0; }
        // PCL: "Leaf size is too small for the input dataset" -> output = *input
        if (out_capacity < n) { set_error("b2vf_filter: output capacity %zu < %zu", out_capacity, n); return B2_ERR_CAPACITY; }
        if (h2d_from == in) pack_cloud_f4(in, n, stride, ioff, h->h_in.as<float>());
        const float *src = h->h_in.as<float>();
        for (size_t j = 0; j < n; ++j) {
            put(j, src + 4 * j);
            if (out_idx) out_idx[j] = -1;
            if (out_cnt) out_cnt[j] = 1;
        }
        *m = n;
        return 0;
    }
    if (M > out_capacity) { set_error("b2vf_filter: output capacity %zu < %zu voxels", out_capacity, M); return B2_ERR_CAPACITY; }
    if ((rc = h->h_out.reserve(M * 16 + 16))) return rc;
    if ((rc = h->h_idx.reserve(M * 4 + 16))) return rc;
    if ((rc = h->h_cnt.reserve(M * 4 + 16))) return rc;
    if (M) {
        B2_CUDA(cudaMemcpyAsync(h->h_out.p, h->d_out.p, M * 16, cudaMemcpyDeviceToHost, h->st));
        if (out_idx) B2_CUDA(cudaMemcpyAsync(h->h_idx.p, h->d_idx.p, M * 4, cudaMemcpyDeviceToHost, h->st));
        if (out_cnt) B2_CUDA(cudaMemcpyAsync(h->h_cnt.p, h->d_cnt.p, M * 4, cudaMemcpyDeviceToHost, h->st));
        B2_CUDA(cudaStreamSynchronize(h->st));
        const float *res = h->h_out.as<float>();
        if (out_stride == 16 && out_ioff == 12) memcpy(out, res, M * 16);
        else for (size_t j = 0; j < M; ++j) put(j, res + 4 * j);
        if (out_idx) memcpy(out_idx, h->h_idx.p, M * 4);
        if (out_cnt) memcpy(out_cnt, h->h_cnt.p, M * 4);
    }
    *m = M;
    return 0;
}

extern "C" int b2cloud_remove_nan(b2cloud *src, b2cloud *dst);

extern "C" int b2vf_filter_batch_device(b2vf *h, const void *d_in_f4, size_t n_total, const uint32_t *h_offsets, size_t B,
                                        void *d_out_f4, uint32_t *d_out_offsets) {
    if (!h || !h_offsets || !d_out_offsets) { set_error("b2vf_filter_batch_device: NULL argument"); return B2_ERR_INVALID; }
    if (B == 0) return 0;
    if (h_offsets[B] != n_total) { set_error("b2vf_filter_batch_device: offsets[B] != n_total"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(h->device));
    int rc;
    if ((rc = h->pipe.plan(h_offsets, B, h->st))) return rc;
    if ((rc = h->pipe.run((const float4 *)d_in_f4, h->leaf[0], h->leaf[1], h->leaf[2], 0, h->st))) return rc;
    if ((rc = vf_launch_centroids(h, (const float4 *)d_in_f4, (float4 *)d_out_f4, nullptr, nullptr, n_total))) return rc;
    B2_CUDA(cudaMemcpyAsync(d_out_offsets, h->pipe.run_seg_off(), (B + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->st));
    return 0;
}

// Batched ingest: like b2vf_filter_batch_device, but the filtered clouds are APPENDED to d_out_f4 behind the ones
// earlier calls left there, without a host round trip: d_cursor[0] (device) counts the points in d_out_f4 so far
// (zero it before the first call), d_cursor[1] becomes 1 when a batch did not fit into out_capacity points (that
// batch is dropped as a whole).  d_out_offsets[first_index + s] receives where cloud s of this batch starts,
// s = 0..B (the last entry is the new cursor, i.e. the first entry of the next batch): after K calls the array is the
// (sum of B) + 1 offsets table b2ndt_align_batch_device takes.
extern "C" int b2vf_filter_batch_append_device(b2vf *h, const void *d_in_f4, size_t n_total, const uint32_t *h_offsets, size_t B,
                                               void *d_out_f4, size_t out_capacity, uint32_t *d_out_offsets, size_t first_index,
                                               uint32_t *d_cursor) {
    if (!h || !h_offsets || !d_out_offsets || !d_cursor || !d_out_f4) { set_error("b2vf_filter_batch_append_device: NULL argument"); return B2_ERR_INVALID; }
    if (B == 0) return 0;
    if (h_offsets[B] != n_total) { set_error("b2vf_filter_batch_append_device: offsets[B] != n_total"); return B2_ERR_INVALID; }
    if (out_capacity >= B2_MAX_POINTS) out_capacity = B2_MAX_POINTS;
    B2_CUDA(cudaSetDevice(h->device));
    int rc;
    if ((rc = h->pipe.plan(h_offsets, B, h->st))) return rc;
    if ((rc = h->pipe.run((const float4 *)d_in_f4, h->leaf[0], h->leaf[1], h->leaf[2], 0, h->st))) return rc;
    if ((rc = vf_launch_centroids(h, (const float4 *)d_in_f4, (float4 *)d_out_f4, nullptr, nullptr, n_total, d_cursor, (uint32_t)out_capacity))) return rc;
    vf_append_kernel<<<1, 256, 0, h->st>>>(h->pipe.run_seg_off(), (uint32_t)B, d_out_offsets + first_index, d_cursor, (uint32_t)out_capacity);
    B2_LAUNCH_CHECK();
    return 0;
}

// Fused ingest + VoxelFilter for a raw scan in HBM: scan de-skew (DistortionAdjust::AdjustCloud, data_pretreat;
// linear_velocity == angular_velocity == NULL skips it) + pcl::removeNaNFromPointCloud (front_end.cpp:92) +
// VoxelFilter::Filter (front_end.cpp:106-107) in ONE pass over the raw points in front of the sort: the de-skewed
// point (or a NaN point for a dropped one) and the bounding box come out of the same kernel.  `filtered` receives
// exactly VoxelFilter(removeNaN(deskew(src))); `ingested` (may be NULL, must not be src) receives the de-skewed cloud
// with src's size, dropped points as NaN points (every consumer of a device cloud skips those; b2cloud_remove_nan
// compacts them away when the count matters).
extern "C" int b2vf_ingest_filter_cloud(b2vf *h, b2cloud *src, float scan_period, const double linear_velocity[3],
                                        const double angular_velocity[3], b2cloud *filtered, b2cloud *ingested) {
    if (!h || !src || !filtered) { set_error("b2vf_ingest_filter_cloud: NULL argument"); return B2_ERR_INVALID; }
    if ((linear_velocity == nullptr) != (angular_velocity == nullptr)) { set_error("b2vf_ingest_filter_cloud: give both velocities or neither"); return B2_ERR_INVALID; }
    if (filtered == src || ingested == src || (ingested && ingested == filtered)) { set_error("b2vf_ingest_filter_cloud: outputs must be distinct from the source and from each other"); return B2_ERR_INVALID; }
    if (src->device != h->device || filtered->device != h->device || (ingested && ingested->device != h->device)) {
        set_error("b2vf_ingest_filter_cloud: handle and clouds live on different devices"); return B2_ERR_INVALID;
    }
    B2_CUDA(cudaSetDevice(h->device));
    const size_t n = src->n;
    filtered->n = 0; filtered->first_known = false;
    if (ingested) { ingested->n = 0; ingested->first_known = false; }
    if (n == 0) return 0;
    int rc;
    IngestOp op;
    memset(&op, 0, sizeof(op));
    if (linear_velocity) {
        float x0 = src->first_xy[0], y0 = src->first_xy[1];
        if (!src->first_known) {       // a cloud produced on the device: fetch its first point (the start azimuth)
            if ((rc = h->h_misc.reserve(256))) return rc;
            B2_CUDA(cudaMemcpyAsync(h->h_misc.p, src->pts.p, 16, cudaMemcpyDeviceToHost, h->st));
            B2_CUDA(cudaStreamSynchronize(h->st));
            x0 = h->h_misc.as<float>()[0]; y0 = h->h_misc.as<float>()[1];
        }
        make_deskew_arg(x0, y0, scan_period, linear_velocity, angular_velocity, op.D);
        op.deskew = 1;
    }
    float4 *ing;
    if (ingested) { if ((rc = ingested->reserve(n))) return rc; ing = ingested->d(); }
    else { if ((rc = h->d_in.reserve(n * 16))) return rc; ing = h->d_in.as<float4>(); }
    if ((rc = filtered->reserve(n))) return rc;
    if ((rc = h->h_misc.reserve(256))) return rc;
    uint32_t off[2] = {0u, (uint32_t)n};
    if ((rc = h->pipe.plan(off, 1, h->st))) return rc;
    if ((rc = h->pipe.run_ingest(src->d(), ing, op, h->leaf[0], h->leaf[1], h->leaf[2], h->st))) return rc;
    if ((rc = vf_launch_centroids(h, ing, filtered->d(), nullptr, nullptr, n))) return rc;
    uint32_t *misc = h->h_misc.as<uint32_t>();
    B2_CUDA(cudaMemcpyAsync(misc, h->pipe.scalars(), 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaMemcpyAsync(misc + 8, h->pipe.layouts(), sizeof(VoxLayout), cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaStreamSynchronize(h->st));
    VoxLayout L;
    memcpy(&L, misc + 8, sizeof(L));
    if (ingested) ingested->n = n;
    if (!L.ok) {
        if (L.n_finite == 0) return 0;
        // PCL: "Leaf size is too small for the input dataset" -> output = *input (here: the kept points, compacted)
        if (!ingested) { set_error("b2vf_ingest_filter_cloud: leaf size too small for the cloud (PCL would return the input)"); return B2_ERR_INVALID; }
        return b2cloud_remove_nan(ingested, filtered);
    }
    filtered->n = misc[1];
    return 0;
}

// Device-resident VoxelFilter::Filter: src and dst are clouds in HBM (src == dst allowed, as the reference
// calls Filter in place: matching.cpp:158, viewer.cpp:207).  Same results as b2vf_filter.
extern "C" int b2vf_filter_cloud(b2vf *h, b2cloud *src, b2cloud *dst) {
    if (!h || !src || !dst) { set_error("b2vf_filter_cloud: NULL argument"); return B2_ERR_INVALID; }
    if (src->device != h->device || dst->device != h->device) { set_error("b2vf_filter_cloud: handle and clouds live on different devices"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(h->device));
    const size_t n = src->n;
    if (n == 0) { dst->n = 0; return 0; }
    int rc;
    const bool in_place = (dst == src);
    float4 *out;
    if (in_place) {
        if ((rc = h->d_out.reserve(n * 16))) return rc;
        out = h->d_out.as<float4>();
    } else {
        dst->n = 0;
        if ((rc = dst->reserve(n))) return rc;
        out = dst->d();
    }
    if ((rc = h->h_misc.reserve(256))) return rc;
    uint32_t off[2] = {0u, (uint32_t)n};
    if ((rc = h->pipe.plan(off, 1, h->st))) return rc;
    if ((rc = h->pipe.run(src->d(), h->leaf[0], h->leaf[1], h->leaf[2], 0, h->st))) return rc;
    if ((rc = vf_launch_centroids(h, src->d(), out, nullptr, nullptr, n))) return rc;
    uint32_t *misc = h->h_misc.as<uint32_t>();
    B2_CUDA(cudaMemcpyAsync(misc, h->pipe.scalars(), 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaMemcpyAsync(misc + 8, h->pipe.layouts(), sizeof(VoxLayout), cudaMemcpyDeviceToHost, h->st));
    B2_CUDA(cudaStreamSynchronize(h->st));
    VoxLayout L;
    memcpy(&L, misc + 8, sizeof(L));
    if (!L.ok) {
        if (L.n_finite == 0) { dst->n = 0; return 0; }
        // PCL: "Leaf size is too small for the input dataset" -> output = *input
        if (!in_place) {
            B2_CUDA(cudaMemcpyAsync(dst->pts.p, src->pts.p, n * 16, cudaMemcpyDeviceToDevice, h->st));
            B2_CUDA(cudaStreamSynchronize(h->st));
        }
        dst->n = n; dst->first_known = false;
        return 0;
    }
    const size_t M = misc[1];
    if (in_place && M) {
        B2_CUDA(cudaMemcpyAsync(src->pts.p, out, M * 16, cudaMemcpyDeviceToDevice, h->st));
        B2_CUDA(cudaStreamSynchronize(h->st));
    }
    dst->n = M; dst->first_known = false;
    return 0;
}
