// b2_common.cuh -- shared host/device helpers for libb2ndt.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>

#include "b2ndt.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb2ndt is written for sm_100a (B200) only"
#endif

namespace b2 {

// ------------------------------------------------------------------ error plumbing ----------
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define B2_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            b2::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return B2_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define B2_LAUNCH_CHECK()                                                                     \
    do {                                                                                      \
        b2::count_launch();                                                                   \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            b2::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return B2_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

// Programmatic dependent launch for chains of small dependent kernels (voxel pipeline, target build): a kernel launched
// through launch_chain() may have its CTAs placed while the kernel before it in the stream is still draining; every such
// kernel starts with chain_sync() -- griddepcontrol.wait returns once the previous kernel has completed and its memory is
// visible -- BEFORE its first global-memory access, so only the launch latency overlaps, never the data.  (In a kernel
// launched the ordinary way the two instructions do nothing.)  B2_PDL=0 in the environment launches the ordinary way.
__device__ __forceinline__ void chain_sync() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
inline bool chain_enabled() {
    static const bool on = [] { const char *e = getenv("B2_PDL"); return !(e && atoi(e) == 0); }();
    return on;
}
template <typename... KArgs, typename... Args>
inline void launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = chain_enabled() ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
    if (e != cudaSuccess && cfg.numAttrs) {        // an attribute the driver refuses: launch the ordinary way
        cudaGetLastError();
        cfg.numAttrs = 0;
        e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
    }
    (void)e;                                       // an error left standing is picked up by B2_LAUNCH_CHECK (cudaGetLastError)
}

// growable device / pinned buffers (never shrink; reused across calls).  Device memory comes from the device's
// stream-ordered pool (cudaMallocAsync) with an unlimited release threshold: a buffer that has to grow -- the local
// map of the front end grows for its first 20 key frames -- gets its new block from the pool instead of paying a
// cudaFree / cudaMalloc pair (tens of milliseconds per key frame over the ~15 buffers of a target build).
inline void devbuf_pool_setup() {
    static thread_local int done_dev = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return; }
    if (dev == done_dev) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    } else {
        cudaGetLastError();
    }
    done_dev = dev;
}
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    // contents are NOT preserved when the buffer grows
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        devbuf_pool_setup();
        release();
        size_t want = bytes + bytes / 2 + 256;
        cudaError_t e = cudaMallocAsync(&p, want, cudaStreamPerThread);
        if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);     // usable from every stream from here on
        if (e != cudaSuccess) { set_error("cudaMallocAsync(%zu) failed: %s", want, cudaGetErrorString(e)); p = nullptr; return B2_ERR_CUDA; }
        cap = want;
        return 0;
    }
    void release() {
        if (p) {
            cudaDeviceSynchronize();                       // work on other streams may still use the block (as cudaFree would wait)
            cudaFreeAsync(p, cudaStreamPerThread);
        }
        p = nullptr; cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) { set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e)); p = nullptr; return B2_ERR_CUDA; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// ------------------------------------------------------------------ voxel layout ------------
// pcl::VoxelGrid / VoxelGridCovariance index layout, one per cloud.
struct VoxLayout {
    int32_t ok;          // 0: no finite point or int32 guard tripped
    int32_t nbits;       // bits needed to sort keys in [0, ncells]  (ncells = invalid-point key)
    float   inv[3];
    float   min_p[3], max_p[3];
    int32_t min_b[3], div_b[3], mul[3];
    uint32_t ncells;     // div_x*div_y*div_z
    uint32_t n_finite;
};

// float <-> order-preserving uint for atomic min/max
__host__ __device__ inline uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ord2f(uint32_t u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

#ifdef __CUDACC__
__device__ __forceinline__ bool finite3(float x, float y, float z) {
    return isfinite(x) && isfinite(y) && isfinite(z);
}
// PCL: ijk = (int)(floor(x * inv) - (float)min_b)  -- float multiply, no FMA contraction
__device__ __forceinline__ int32_t vox_index(const VoxLayout &L, float x, float y, float z) {
    int i0 = (int)(floorf(__fmul_rn(x, L.inv[0])) - (float)L.min_b[0]);
    int i1 = (int)(floorf(__fmul_rn(y, L.inv[1])) - (float)L.min_b[1]);
    int i2 = (int)(floorf(__fmul_rn(z, L.inv[2])) - (float)L.min_b[2]);
    return i0 * L.mul[0] + i1 * L.mul[1] + i2 * L.mul[2];
}
// pcl::transformPointCloud, float, evaluated left to right without contraction
__device__ __forceinline__ void transform_f32(const float *T /*col-major 4x4*/, float x, float y, float z,
                                              float &ox, float &oy, float &oz) {
    ox = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], x), __fmul_rn(T[4], y)), __fmul_rn(T[8], z)), T[12]);
    oy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[1], x), __fmul_rn(T[5], y)), __fmul_rn(T[9], z)), T[13]);
    oz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[2], x), __fmul_rn(T[6], y)), __fmul_rn(T[10], z)), T[14]);
}
#endif

#ifdef __CUDACC__
// 16-byte asynchronous copy global -> shared (LDGSTS), grouped: the gathers of the crowded-voxel kernels go straight
// into the staging ring without holding a register (and the issuing warp) until they arrive
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

// ------------------------------------------------------------------ sort plan ----------------
// A "tile" is a run of at most tile_elems consecutive elements of ONE cloud (segment) of a concatenated
// batch; kernels are launched one CTA per tile.  Two tile sizes: 256 threads x 4 keys for small inputs (enough
// CTAs to fill 148 SMs from ~1e5 points), 256 x 16 for large ones (less look-back work per key).
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS_SMALL = 4, SORT_ITEMS_LARGE = 16;
constexpr size_t SORT_LARGE_FROM = 600000;         // points in the batch from which the large tile is used (>= 147 tiles)
constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;

struct TileDesc {
    uint32_t seg;        // cloud index
    uint32_t begin;      // first element (global index into the concatenated arrays)
    uint32_t count;      // elements in this tile
    uint32_t tile_in_seg;
};
struct SegDesc {
    uint32_t begin, count;      // element range
    uint32_t tile_begin, ntiles;
};

}  // namespace b2
