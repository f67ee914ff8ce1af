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
    std::vector<BB> bbs(nt);
    if (nt <= 1) work(0, n, &bbs[0]);
    else {
        std::vector<std::thread> th;
        size_t chunk = (n + nt - 1) / nt;
        for (size_t t = 0; t < nt; ++t) {
            size_t a = t * chunk, b = a + chunk < n ? a + chunk : n;
            if (a >= b) { for (int k = 0; k < 3; ++k) { bbs[t].mn[k] = FLT_MAX; bbs[t].mx[k] = -FLT_MAX; } continue; }
            th.emplace_back(work, a, b, &bbs[t]);
        }
        for (auto &t : th) t.join();
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
    std::vector<std::thread> th;
    size_t chunk = (n + nt - 1) / nt;
    for (size_t t = 0; t < nt; ++t) {
        size_t a = t * chunk, b = a + chunk < n ? a + chunk : n;
        if (a < b) th.emplace_back(work, a, b);
    }
    for (auto &t : th) t.join();
}

#ifndef B2_SCATTER_MIN_BLOCKS
#define B2_SCATTER_MIN_BLOCKS 4      // resident CTAs per SM the scatter kernel is compiled for (register cap 64): 5 M-point filter 0.65 -> 0.53 ms
#endif
// ------------------------------------------------------------------ kernels -----------------
__global__ void bbox_init_kernel(uint32_t *bbox, uint32_t *scalars, uint32_t B) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) {
        uint32_t *b = bbox + 8 * i;
        b[0] = b[1] = b[2] = 0xFFFFFFFFu;
        b[3] = b[4] = b[5] = 0u;
        b[6] = 0u; b[7] = 0u;
    }
    if (i < 8) scalars[i] = 0u;
}

__global__ void __launch_bounds__(SORT_THREADS) bbox_kernel(const float4 *__restrict__ pts,
                                                            const TileDesc *__restrict__ tiles,
                                                            uint32_t *__restrict__ bbox) {
    const TileDesc t = tiles[blockIdx.x];
    float mn0 = FLT_MAX, mn1 = FLT_MAX, mn2 = FLT_MAX, mx0 = -FLT_MAX, mx1 = -FLT_MAX, mx2 = -FLT_MAX;
    uint32_t cnt = 0;
    for (uint32_t k = threadIdx.x; k < t.count; k += SORT_THREADS) {
        float4 p = __ldg(&pts[t.begin + k]);
        if (finite3(p.x, p.y, p.z)) {
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
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s[w][0] = mn0; s[w][1] = mn1; s[w][2] = mn2; s[w][3] = mx0; s[w][4] = mx1; s[w][5] = mx2; sc[w] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t c = 0;
        for (int i = 0; i < SORT_THREADS / 32; ++i) {
            c += sc[i];
            mn0 = fminf(mn0, s[i][0]); mn1 = fminf(mn1, s[i][1]); mn2 = fminf(mn2, s[i][2]);
            mx0 = fmaxf(mx0, s[i][3]); mx1 = fmaxf(mx1, s[i][4]); mx2 = fmaxf(mx2, s[i][5]);
        }
        if (c) {
            uint32_t *b = bbox + 8 * t.seg;
            atomicMin(&b[0], f2ord(mn0)); atomicMin(&b[1], f2ord(mn1)); atomicMin(&b[2], f2ord(mn2));
            atomicMax(&b[3], f2ord(mx0)); atomicMax(&b[4], f2ord(mx1)); atomicMax(&b[5], f2ord(mx2));
            atomicAdd(&b[6], c);
        }
    }
}

// pcl::VoxelGrid::applyFilter prologue: inverse leaf, int64 overflow guard, min_b / div_b / divb_mul
__global__ void layout_kernel(const uint32_t *__restrict__ bbox, float lx, float ly, float lz,
                              VoxLayout *__restrict__ layouts, uint32_t *__restrict__ scalars, uint32_t B) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B) return;
    const uint32_t *b = bbox + 8 * s;
    VoxLayout L;
    memset(&L, 0, sizeof(L));
    const float leaf[3] = {lx, ly, lz};
    for (int a = 0; a < 3; ++a) L.inv[a] = __fdiv_rn(1.0f, leaf[a]);
    L.n_finite = b[6];
    if (L.n_finite > 0) {
        long long d[3];
        for (int a = 0; a < 3; ++a) {
            L.min_p[a] = ord2f(b[a]);
            L.max_p[a] = ord2f(b[3 + a]);
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
    if (L.ok) {
        L.nbits = 32 - __clz(L.ncells);   // keys lie in [0, ncells] (ncells = non-finite points)
        atomicMax(&scalars[0], (uint32_t)L.nbits);
    }
    layouts[s] = L;
}

__global__ void __launch_bounds__(SORT_THREADS) key_kernel(const float4 *__restrict__ pts,
                                                           const TileDesc *__restrict__ tiles,
                                                           const VoxLayout *__restrict__ layouts,
                                                           uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const TileDesc t = tiles[blockIdx.x];
    __shared__ VoxLayout L;
    if (threadIdx.x == 0) L = layouts[t.seg];
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < t.count; k += SORT_THREADS) {
        uint32_t i = t.begin + k;
        float4 p = __ldg(&pts[i]);
        uint32_t key = L.ncells;
        if (L.ok && finite3(p.x, p.y, p.z)) key = (uint32_t)vox_index(L, p.x, p.y, p.z);
        keys[i] = key;
        vals[i] = i;
    }
}

// ---- radix pass: per-tile digit histogram -> [seg][bin][tile] table
__global__ void __launch_bounds__(SORT_THREADS) radix_hist_kernel(const uint32_t *__restrict__ keys,
                                                                  const TileDesc *__restrict__ tiles,
                                                                  const SegDesc *__restrict__ segs,
                                                                  uint32_t *__restrict__ tilehist, int shift,
                                                                  const uint32_t *__restrict__ scalars) {
    if ((uint32_t)shift >= scalars[0]) return;
    const TileDesc t = tiles[blockIdx.x];
    const SegDesc sg = segs[t.seg];
    __shared__ uint32_t h[RADIX];
    h[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < t.count; k += SORT_THREADS)
        atomicAdd(&h[(__ldg(&keys[t.begin + k]) >> shift) & (RADIX - 1)], 1u);
    __syncthreads();
    tilehist[(size_t)sg.tile_begin * RADIX + (size_t)threadIdx.x * sg.ntiles + t.tile_in_seg] = h[threadIdx.x];
}

// ---- radix pass: exclusive scan of every (cloud, bin) row over the cloud's tiles.  One warp per row,
// eight rows per CTA, 32 CTAs per cloud (a single CTA per cloud took 330 us per pass on a 5 M-point map).
// The row totals go to binbase[cloud][bin]; the scatter kernel turns them into bin offsets itself.
__global__ void __launch_bounds__(SORT_THREADS) radix_scan_kernel(const SegDesc *__restrict__ segs,
                                                                  uint32_t *__restrict__ tilehist,
                                                                  uint32_t *__restrict__ binbase, int shift,
                                                                  const uint32_t *__restrict__ scalars) {
    if ((uint32_t)shift >= scalars[0]) return;
    constexpr int ROWS_PER_CTA = SORT_THREADS / 32;
    const uint32_t seg = blockIdx.x / (RADIX / ROWS_PER_CTA);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int bin = (blockIdx.x % (RADIX / ROWS_PER_CTA)) * ROWS_PER_CTA + w;
    const SegDesc sg = segs[seg];
    uint32_t *row = tilehist + (size_t)sg.tile_begin * RADIX + (size_t)bin * sg.ntiles;
    uint32_t running = 0;
    for (uint32_t t0 = 0; t0 < sg.ntiles; t0 += 32) {
        uint32_t v = (t0 + l < sg.ntiles) ? row[t0 + l] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
            if (l >= o) inc += n;
        }
        if (t0 + l < sg.ntiles) row[t0 + l] = running + inc - v;
        running += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (l == 0) binbase[(size_t)seg * RADIX + bin] = running;      // row total
}

// ---- radix pass: stable scatter.  Warp w owns elements [512w, 512w+512) of the tile in 16 rounds of
// 32 consecutive keys; ranks come from match_any + per-warp digit counters in shared memory.
__global__ void __launch_bounds__(SORT_THREADS, B2_SCATTER_MIN_BLOCKS) radix_scatter_kernel(
    const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ keys_out,
    uint32_t *__restrict__ vals_out, const TileDesc *__restrict__ tiles, const SegDesc *__restrict__ segs,
    const uint32_t *__restrict__ tilehist, const uint32_t *__restrict__ binbase, int shift,
    const uint32_t *__restrict__ scalars) {
    const TileDesc t = tiles[blockIdx.x];
    if ((uint32_t)shift >= scalars[0]) {   // digit is zero everywhere: identity pass
        for (uint32_t k = threadIdx.x; k < t.count; k += SORT_THREADS) {
            keys_out[t.begin + k] = keys_in[t.begin + k];
            vals_out[t.begin + k] = vals_in[t.begin + k];
        }
        return;
    }
    const SegDesc sg = segs[t.seg];
    constexpr int NW = SORT_THREADS / 32;
    constexpr int ROUNDS = SORT_TILE / SORT_THREADS;
    __shared__ uint32_t wcnt[NW][RADIX];
    __shared__ uint32_t gbase[RADIX];
    __shared__ uint32_t wtot[NW];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    for (int i = threadIdx.x; i < NW * RADIX; i += SORT_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    uint32_t key[ROUNDS], val[ROUNDS];
    uint16_t loc[ROUNDS];
    const uint32_t lt = (1u << l) - 1u;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        uint32_t e = w * (32 * ROUNDS) + r * 32 + l;
        bool valid = e < t.count;
        key[r] = valid ? keys_in[t.begin + e] : 0u;
        val[r] = valid ? vals_in[t.begin + e] : 0u;
        uint32_t digit = valid ? ((key[r] >> shift) & (RADIX - 1)) : (uint32_t)RADIX;
        uint32_t peers = __match_any_sync(0xffffffffu, digit);
        int leader = __ffs(peers) - 1;
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
    {
        uint32_t d = threadIdx.x, run = 0;
#pragma unroll
        for (int i = 0; i < NW; ++i) { uint32_t c = wcnt[i][d]; wcnt[i][d] = run; run += c; }
        // exclusive scan of the cloud's 256 bin totals (block scan)
        const uint32_t v = binbase[(size_t)t.seg * RADIX + d];
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
            if (l >= o) inc += n;
        }
        if (l == 31) wtot[w] = inc;
        __syncthreads();
        uint32_t add = 0;
        for (int i = 0; i < w; ++i) add += wtot[i];
        gbase[d] = sg.begin + (add + inc - v) +
                   tilehist[(size_t)sg.tile_begin * RADIX + (size_t)d * sg.ntiles + t.tile_in_seg];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        uint32_t e = w * (32 * ROUNDS) + r * 32 + l;
        if (e < t.count) {
            uint32_t digit = (key[r] >> shift) & (RADIX - 1);
            uint32_t pos = gbase[digit] + wcnt[w][digit] + loc[r];
            keys_out[pos] = key[r];
            vals_out[pos] = val[r];
        }
    }
}

// ---- run heads (one run = one occupied voxel).  Blocked arrangement: thread t owns 16 consecutive keys.
template <bool WRITE>
__global__ void __launch_bounds__(SORT_THREADS) head_kernel(const uint32_t *__restrict__ keys,
                                                            const TileDesc *__restrict__ tiles,
                                                            const SegDesc *__restrict__ segs,
                                                            const VoxLayout *__restrict__ layouts,
                                                            uint32_t *__restrict__ tile_heads,
                                                            uint32_t *__restrict__ run_start,
                                                            uint32_t *__restrict__ run_seg) {
    const TileDesc t = tiles[blockIdx.x];
    const SegDesc sg = segs[t.seg];
    __shared__ uint32_t s_ok, s_inv;
    __shared__ uint32_t wsum[SORT_THREADS / 32];
    if (threadIdx.x == 0) { s_ok = layouts[t.seg].ok; s_inv = layouts[t.seg].ncells; }
    __syncthreads();
    constexpr int PER = SORT_TILE / SORT_THREADS;
    const uint32_t e0 = threadIdx.x * PER;
    uint32_t flags = 0;
    if (s_ok && e0 < t.count) {
        uint32_t gi = t.begin + e0;
        uint32_t prev = (gi > sg.begin) ? keys[gi - 1] : 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (e0 + j < t.count) {
                uint32_t k = keys[gi + j];
                bool head = (k != s_inv) && ((gi + j == sg.begin) || (k != prev));
                flags |= head ? (1u << j) : 0u;
                prev = k;
            }
        }
    }
    uint32_t cnt = __popc(flags), inc = cnt;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (l >= o) inc += n;
    }
    if (l == 31) wsum[w] = inc;
    __syncthreads();
    uint32_t add = 0;
    for (int i = 0; i < w; ++i) add += wsum[i];
    if (!WRITE) {
        if (threadIdx.x == SORT_THREADS - 1) tile_heads[blockIdx.x] = add + inc;
    } else {
        uint32_t o = tile_heads[blockIdx.x] + add + inc - cnt;
#pragma unroll
        for (int j = 0; j < PER; ++j)
            if (flags & (1u << j)) { run_start[o] = t.begin + e0 + j; run_seg[o] = t.seg; ++o; }
    }
}

// single-CTA exclusive scan of the per-tile head counts (+ per-cloud first-run table)
__global__ void __launch_bounds__(1024) scan_tiles_kernel(uint32_t *__restrict__ tile_heads, uint32_t ntiles,
                                                          const SegDesc *__restrict__ segs, uint32_t B,
                                                          uint32_t *__restrict__ run_seg_off,
                                                          uint32_t *__restrict__ scalars) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < ntiles; t0 += 1024) {
        uint32_t i = t0 + threadIdx.x;
        uint32_t v = i < ntiles ? tile_heads[i] : 0u, inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
            if (l >= o) inc += n;
        }
        if (l == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            uint32_t x = wsum[l], xi = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t n = __shfl_up_sync(0xffffffffu, xi, o);
                if (l >= o) xi += n;
            }
            wsum[l] = xi - x;
        }
        __syncthreads();
        uint32_t excl = carry + wsum[w] + inc - v;
        if (i < ntiles) tile_heads[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) { tile_heads[ntiles] = carry; scalars[1] = carry; }
    __syncthreads();
    for (uint32_t s = threadIdx.x; s <= B; s += 1024) {
        uint32_t tb = (s < B) ? segs[s].tile_begin : ntiles;
        run_seg_off[s] = tile_heads[tb];
    }
}

// ------------------------------------------------------------------ VoxPipeline -------------
int VoxPipeline::plan(const uint32_t *off, size_t nB, cudaStream_t st) {
    B = nB;
    N = off[nB];
    h_tiles.clear();
    h_segs.resize(nB);
    for (size_t s = 0; s < nB; ++s) {
        uint32_t b = off[s], c = off[s + 1] - off[s];
        SegDesc sd;
        sd.begin = b; sd.count = c; sd.tile_begin = (uint32_t)h_tiles.size();
        sd.ntiles = (c + SORT_TILE - 1) / SORT_TILE;
        for (uint32_t t = 0; t < sd.ntiles; ++t) {
            TileDesc td;
            td.seg = (uint32_t)s; td.begin = b + t * SORT_TILE;
            td.count = (c - t * SORT_TILE < (uint32_t)SORT_TILE) ? c - t * SORT_TILE : SORT_TILE;
            td.tile_in_seg = t;
            h_tiles.push_back(td);
        }
        h_segs[s] = sd;
    }
    ntiles = h_tiles.size();
    int rc;
    if ((rc = d_tiles.reserve((ntiles + 1) * sizeof(TileDesc)))) return rc;
    if ((rc = d_segs.reserve((B + 1) * sizeof(SegDesc)))) return rc;
    if ((rc = d_bbox.reserve((B + 1) * 8 * sizeof(uint32_t)))) return rc;
    if ((rc = d_layouts.reserve((B + 1) * sizeof(VoxLayout)))) return rc;
    if ((rc = d_scalars.reserve(8 * sizeof(uint32_t)))) return rc;
    for (int i = 0; i < 2; ++i) {
        if ((rc = d_keys[i].reserve((N + 1) * sizeof(uint32_t)))) return rc;
        if ((rc = d_vals[i].reserve((N + 1) * sizeof(uint32_t)))) return rc;
    }
    if ((rc = d_tilehist.reserve((ntiles + 1) * RADIX * sizeof(uint32_t)))) return rc;
    if ((rc = d_binbase.reserve((B + 1) * RADIX * sizeof(uint32_t)))) return rc;
    if ((rc = d_tile_heads.reserve((ntiles + 2) * sizeof(uint32_t)))) return rc;
    if ((rc = d_run_start.reserve((N + 2) * sizeof(uint32_t)))) return rc;
    if ((rc = d_run_seg.reserve((N + 2) * sizeof(uint32_t)))) return rc;
    if ((rc = d_run_seg_off.reserve((B + 2) * sizeof(uint32_t)))) return rc;
    if (ntiles) B2_CUDA(cudaMemcpyAsync(d_tiles.p, h_tiles.data(), ntiles * sizeof(TileDesc), cudaMemcpyHostToDevice, st));
    if (B) B2_CUDA(cudaMemcpyAsync(d_segs.p, h_segs.data(), B * sizeof(SegDesc), cudaMemcpyHostToDevice, st));
    // the async copies read h_tiles/h_segs (pageable): the runtime stages them before returning
    return 0;
}

int VoxPipeline::run(const float4 *d_pts, float lx, float ly, float lz, int nbits_hint, cudaStream_t st) {
    const uint32_t nB = (uint32_t)B, nt = (uint32_t)ntiles;
    uint32_t *bbox = d_bbox.as<uint32_t>(), *sc = d_scalars.as<uint32_t>();
    const TileDesc *tiles = d_tiles.as<TileDesc>();
    const SegDesc *segs = d_segs.as<SegDesc>();
    VoxLayout *lay = d_layouts.as<VoxLayout>();
    bbox_init_kernel<<<(nB + 8 + 255) / 256, 256, 0, st>>>(bbox, sc, nB);
    B2_LAUNCH_CHECK();
    if (nt) {
        bbox_kernel<<<nt, SORT_THREADS, 0, st>>>(d_pts, tiles, bbox);
        B2_LAUNCH_CHECK();
    }
    layout_kernel<<<(nB + 127) / 128, 128, 0, st>>>(bbox, lx, ly, lz, lay, sc, nB);
    B2_LAUNCH_CHECK();
    final_buf = 0;
    if (nt) {
        key_kernel<<<nt, SORT_THREADS, 0, st>>>(d_pts, tiles, lay, d_keys[0].as<uint32_t>(), d_vals[0].as<uint32_t>());
        B2_LAUNCH_CHECK();
        int passes = 32 / RADIX_BITS;
        if (nbits_hint > 0) passes = (nbits_hint + RADIX_BITS - 1) / RADIX_BITS;
        for (int p = 0; p < passes; ++p) {
            int shift = p * RADIX_BITS;
            uint32_t *ki = d_keys[final_buf].as<uint32_t>(), *vi = d_vals[final_buf].as<uint32_t>();
            uint32_t *ko = d_keys[final_buf ^ 1].as<uint32_t>(), *vo = d_vals[final_buf ^ 1].as<uint32_t>();
            radix_hist_kernel<<<nt, SORT_THREADS, 0, st>>>(ki, tiles, segs, d_tilehist.as<uint32_t>(), shift, sc);
            B2_LAUNCH_CHECK();
            radix_scan_kernel<<<nB * (RADIX / (SORT_THREADS / 32)), SORT_THREADS, 0, st>>>(segs, d_tilehist.as<uint32_t>(), d_binbase.as<uint32_t>(), shift, sc);
            B2_LAUNCH_CHECK();
            radix_scatter_kernel<<<nt, SORT_THREADS, 0, st>>>(ki, vi, ko, vo, tiles, segs, d_tilehist.as<uint32_t>(),
                                                             d_binbase.as<uint32_t>(), shift, sc);
            B2_LAUNCH_CHECK();
            final_buf ^= 1;
        }
        head_kernel<false><<<nt, SORT_THREADS, 0, st>>>(sorted_keys(), tiles, segs, lay, d_tile_heads.as<uint32_t>(),
                                                       nullptr, nullptr);
        B2_LAUNCH_CHECK();
    }
    scan_tiles_kernel<<<1, 1024, 0, st>>>(d_tile_heads.as<uint32_t>(), nt, segs, nB, d_run_seg_off.as<uint32_t>(), sc);
    B2_LAUNCH_CHECK();
    if (nt) {
        head_kernel<true><<<nt, SORT_THREADS, 0, st>>>(sorted_keys(), tiles, segs, lay, d_tile_heads.as<uint32_t>(),
                                                      d_run_start.as<uint32_t>(), d_run_seg.as<uint32_t>());
        B2_LAUNCH_CHECK();
    }
    return 0;
}

__global__ void prepared_layout_kernel(VoxLayout *layouts, uint32_t *scalars, uint32_t B, uint32_t invalid_key, int nbits) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < B) {
        VoxLayout L;
        memset(&L, 0, sizeof(L));
        L.ok = 1; L.ncells = invalid_key; L.nbits = nbits;
        layouts[s] = L;
    }
    if (s == 0) { scalars[0] = (uint32_t)nbits; scalars[1] = 0u; }
}

int VoxPipeline::run_prepared(uint32_t invalid_key, int nbits, cudaStream_t st) {
    const uint32_t nB = (uint32_t)B, nt = (uint32_t)ntiles;
    uint32_t *sc = d_scalars.as<uint32_t>();
    const TileDesc *tiles = d_tiles.as<TileDesc>();
    const SegDesc *segs = d_segs.as<SegDesc>();
    VoxLayout *lay = d_layouts.as<VoxLayout>();
    prepared_layout_kernel<<<(nB + 127) / 128, 128, 0, st>>>(lay, sc, nB, invalid_key, nbits);
    B2_LAUNCH_CHECK();
    final_buf = 0;
    if (nt) {
        const int passes = (nbits + RADIX_BITS - 1) / RADIX_BITS;
        for (int p = 0; p < passes; ++p) {
            int shift = p * RADIX_BITS;
            uint32_t *ki = d_keys[final_buf].as<uint32_t>(), *vi = d_vals[final_buf].as<uint32_t>();
            uint32_t *ko = d_keys[final_buf ^ 1].as<uint32_t>(), *vo = d_vals[final_buf ^ 1].as<uint32_t>();
            radix_hist_kernel<<<nt, SORT_THREADS, 0, st>>>(ki, tiles, segs, d_tilehist.as<uint32_t>(), shift, sc);
            B2_LAUNCH_CHECK();
            radix_scan_kernel<<<nB * (RADIX / (SORT_THREADS / 32)), SORT_THREADS, 0, st>>>(segs, d_tilehist.as<uint32_t>(), d_binbase.as<uint32_t>(), shift, sc);
            B2_LAUNCH_CHECK();
            radix_scatter_kernel<<<nt, SORT_THREADS, 0, st>>>(ki, vi, ko, vo, tiles, segs, d_tilehist.as<uint32_t>(),
                                                             d_binbase.as<uint32_t>(), shift, sc);
            B2_LAUNCH_CHECK();
            final_buf ^= 1;
        }
        head_kernel<false><<<nt, SORT_THREADS, 0, st>>>(sorted_keys(), tiles, segs, lay, d_tile_heads.as<uint32_t>(), nullptr, nullptr);
        B2_LAUNCH_CHECK();
    }
    scan_tiles_kernel<<<1, 1024, 0, st>>>(d_tile_heads.as<uint32_t>(), nt, segs, nB, d_run_seg_off.as<uint32_t>(), sc);
    B2_LAUNCH_CHECK();
    if (nt) {
        head_kernel<true><<<nt, SORT_THREADS, 0, st>>>(sorted_keys(), tiles, segs, lay, d_tile_heads.as<uint32_t>(),
                                                      d_run_start.as<uint32_t>(), d_run_seg.as<uint32_t>());
        B2_LAUNCH_CHECK();
    }
    return 0;
}

void VoxPipeline::release() {
    d_tiles.release(); d_segs.release(); d_bbox.release(); d_layouts.release(); d_scalars.release();
    for (int i = 0; i < 2; ++i) { d_keys[i].release(); d_vals[i].release(); }
    d_tilehist.release(); d_binbase.release(); d_tile_heads.release();
    d_run_start.release(); d_run_seg.release(); d_run_seg_off.release();
}

// ------------------------------------------------------------------ voxel filter ------------
// One warp per occupied voxel.  Lanes gather 32 member points at a time (in stable, i.e. input,
// order) and every lane then folds them in sequentially through shuffles, so the float centroid is the
// same left-to-right float sum pcl::VoxelGrid forms for that member order.
__global__ void __launch_bounds__(256) vf_centroid_kernel(const float4 *__restrict__ pts,
                                                          const uint32_t *__restrict__ keys,
                                                          const uint32_t *__restrict__ vals,
                                                          const uint32_t *__restrict__ run_start,
                                                          const uint32_t *__restrict__ run_seg,
                                                          const uint32_t *__restrict__ run_seg_off,
                                                          const SegDesc *__restrict__ segs,
                                                          const VoxLayout *__restrict__ layouts,
                                                          const uint32_t *__restrict__ scalars,
                                                          float4 *__restrict__ out, int32_t *__restrict__ out_idx,
                                                          int32_t *__restrict__ out_cnt) {
    const uint32_t total = scalars[1];
    const int l = threadIdx.x & 31;
    for (uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < total; j += (gridDim.x * blockDim.x) >> 5) {
        const uint32_t sg = run_seg[j];
        const uint32_t s = run_start[j];
        const uint32_t e = (j + 1 < run_seg_off[sg + 1]) ? run_start[j + 1] : segs[sg].begin + layouts[sg].n_finite;
        float ax = 0.f, ay = 0.f, az = 0.f, ai = 0.f;
        for (uint32_t c = s; c < e; c += 32) {
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c + l < e) p = __ldg(&pts[vals[c + l]]);
            const int m = (e - c < 32u) ? (int)(e - c) : 32;
            if (m == 32) {
                // full chunk: all 128 shuffles are independent of the four add chains, so the fold costs
                // ~one FADD latency per member (crowded voxels hold thousands of points)
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    ax = __fadd_rn(ax, __shfl_sync(0xffffffffu, p.x, k));
                    ay = __fadd_rn(ay, __shfl_sync(0xffffffffu, p.y, k));
                    az = __fadd_rn(az, __shfl_sync(0xffffffffu, p.z, k));
                    ai = __fadd_rn(ai, __shfl_sync(0xffffffffu, p.w, k));
                }
            } else {
                for (int k = 0; k < m; ++k) {
                    ax = __fadd_rn(ax, __shfl_sync(0xffffffffu, p.x, k));
                    ay = __fadd_rn(ay, __shfl_sync(0xffffffffu, p.y, k));
                    az = __fadd_rn(az, __shfl_sync(0xffffffffu, p.z, k));
                    ai = __fadd_rn(ai, __shfl_sync(0xffffffffu, p.w, k));
                }
            }
        }
        if (l == 0) {
            const float n = (float)(e - s);
            out[j] = make_float4(__fdiv_rn(ax, n), __fdiv_rn(ay, n), __fdiv_rn(az, n), __fdiv_rn(ai, n));
            if (out_idx) out_idx[j] = (int32_t)keys[s];
            if (out_cnt) out_cnt[j] = (int32_t)(e - s);
        }
    }
}

}  // namespace b2

// ------------------------------------------------------------------ C ABI: b2vf_* ------------
using namespace b2;

struct b2vf {
    int device = 0;
    float leaf[3];
    cudaStream_t own = nullptr, st = nullptr;
    VoxPipeline pipe;
    DevBuf d_in, d_out, d_idx, d_cnt;
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
    h->d_in.release(); h->d_out.release(); h->d_idx.release(); h->d_cnt.release();
    h->h_in.release(); h->h_out.release(); h->h_idx.release(); h->h_cnt.release(); h->h_misc.release();
    if (h->own) cudaStreamDestroy(h->own);
    delete h;
}

extern "C" int b2vf_set_stream(b2vf *h, void *stream) {
    if (!h) { set_error("b2vf_set_stream: NULL handle"); return B2_ERR_INVALID; }
    h->st = stream ? (cudaStream_t)stream : h->own;
    return 0;
}

static int vf_launch_centroids(b2vf *h, const float4 *d_in, float4 *d_out, int32_t *d_idx, int32_t *d_cnt, size_t n) {
    if (n == 0) return 0;
    // enough warps for every voxel of a typical cloud, grid-stride beyond that; multiple of 148 SMs
    unsigned blocks = (unsigned)((n / 8 + 7) / 8);
    if (blocks < 148) blocks = 148;
    if (blocks > 148 * 16) blocks = 148 * 16;
    vf_centroid_kernel<<<blocks, 256, 0, h->st>>>(d_in, h->pipe.sorted_keys(), h->pipe.sorted_vals(), h->pipe.run_start(),
                                                 h->pipe.run_seg(), h->pipe.run_seg_off(), h->pipe.d_segs.as<SegDesc>(),
                                                 h->pipe.layouts(), h->pipe.scalars(), d_out, d_idx, d_cnt);
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
    if (n >= 0xFFFFFFF0ull) { set_error("b2vf_filter: cloud too large"); return B2_ERR_INVALID; }
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
        dst->n = n;
        return 0;
    }
    const size_t M = misc[1];
    if (in_place && M) {
        B2_CUDA(cudaMemcpyAsync(src->pts.p, out, M * 16, cudaMemcpyDeviceToDevice, h->st));
        B2_CUDA(cudaStreamSynchronize(h->st));
    }
    dst->n = M;
    return 0;
}
