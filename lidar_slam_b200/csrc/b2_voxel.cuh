// b2_voxel.cuh -- voxelisation pipeline shared by the VoxelFilter, the NDT target-grid build and the height map:
//   bbox + PCL voxel layout (one kernel) -> voxel key per point + digit histograms (one kernel) -> stable LSD radix
//   sort by key, ONE kernel per 8-bit digit (decoupled look-back, "onesweep"), batched over independent clouds ->
//   run heads (one run = one occupied voxel; one kernel, decoupled look-back).
// Replaces pcl::VoxelGrid's std::sort of (idx, point) pairs and VoxelGridCovariance's std::map
// insertion.
#pragma once

#include <vector>

#include "b2_common.cuh"

namespace b2 {

// what the consumers of a finished sort read (device side).  Which ping-pong buffer holds the result depends on
// the number of radix passes that actually ran, and that is decided on the device (scalars[0] = key width in bits).
struct SortView {
    const uint32_t *k[2];
    const uint32_t *v[2];
    const uint32_t *scalars;      // [0] key width in bits (max over clouds), [1] total runs
#ifdef __CUDACC__
    __device__ __forceinline__ int sel() const { return (int)(((__ldg(&scalars[0]) + RADIX_BITS - 1) / RADIX_BITS) & 1u); }
    __device__ __forceinline__ const uint32_t *keys() const { return sel() ? k[1] : k[0]; }
    __device__ __forceinline__ const uint32_t *vals() const { return sel() ? v[1] : v[0]; }
#endif
};

// the tiling of a batch of clouds (device side).  A single cloud needs no descriptor arrays: tiles == nullptr.
struct PlanView {
    const TileDesc *tiles;
    const SegDesc *segs;
    uint32_t N, tile_elems, ntiles, B;
};

struct VoxPipeline {
    // plan (host + device copies; only uploaded for batches of more than one cloud)
    std::vector<TileDesc> h_tiles;
    std::vector<SegDesc> h_segs;
    std::vector<uint32_t> h_off;
    DevBuf d_tiles, d_segs;
    // per segment
    DevBuf d_layouts;   // VoxLayout[B]
    DevBuf d_scalars;   // uint32[16]: [0] max key bits, [1] total runs, [2] CTAs done (bbox kernel), [4..8] tile tickets
                        //             of radix pass 0..3 and of the run-head kernel
    DevBuf d_hist;      // uint32[B][4][256]  digit histograms of every radix pass
    // per element (ping-pong)
    DevBuf d_keys[2], d_vals[2];
    // per tile
    DevBuf d_bbox_part; // float[ntiles][8]    per-tile bounding box + finite count
    DevBuf d_state;     // uint32[4][ntiles][256] look-back state of the radix passes, then uint32[ntiles] of the run heads
    // runs
    DevBuf d_run_start;    // uint32[N+1]   global element index of each run's first element
    DevBuf d_run_seg;      // uint32[N]     cloud of each run
    DevBuf d_run_seg_off;  // uint32[B+1]   first run of each cloud
    size_t N = 0, B = 0, ntiles = 0;
    uint32_t tile_elems = 1024;
    bool scalars_ready = false;

    // Build the plan for B clouds described by host offsets (B+1 entries, in points).
    int plan(const uint32_t *h_offsets, size_t B, cudaStream_t st);
    // Run bbox/layout/keys/sort/run-heads on packed float4 points (device).  nbits_hint > 0 limits the
    // number of radix passes from the host side (must be >= the true key width); 0 = decide on device.
    int run(const float4 *d_pts, float lx, float ly, float lz, int nbits_hint, cudaStream_t st);
    // The same with the fused ingest in front (b2_cloud.cuh): d_raw is transformed / marked point by point into
    // d_ingested by the first kernel, everything after it works on d_ingested.
    int run_ingest(const float4 *d_raw, float4 *d_ingested, const struct IngestOp &op, float lx, float ly, float lz, cudaStream_t st);
    // Generic stable sort + run detection for keys the caller wrote itself: after plan(), fill keys0() (one
    // uint32 key per element, key == invalid_key drops the element to the end / out of the runs) and call this with
    // the key width in bits.  Results as after run().
    int run_prepared(uint32_t invalid_key, int nbits, cudaStream_t st);
    uint32_t *keys0() { return d_keys[0].as<uint32_t>(); }

    SortView view() const {
        SortView s;
        s.k[0] = d_keys[0].as<uint32_t>(); s.k[1] = d_keys[1].as<uint32_t>();
        s.v[0] = d_vals[0].as<uint32_t>(); s.v[1] = d_vals[1].as<uint32_t>();
        s.scalars = d_scalars.as<uint32_t>();
        return s;
    }
    PlanView plan_view() const {
        PlanView p;
        p.tiles = (B > 1) ? d_tiles.as<TileDesc>() : nullptr;
        p.segs = (B > 1) ? d_segs.as<SegDesc>() : nullptr;
        p.N = (uint32_t)N; p.tile_elems = tile_elems; p.ntiles = (uint32_t)ntiles; p.B = (uint32_t)B;
        return p;
    }
    const VoxLayout *layouts() const { return d_layouts.as<VoxLayout>(); }
    const uint32_t *run_start() const { return d_run_start.as<uint32_t>(); }
    const uint32_t *run_seg() const { return d_run_seg.as<uint32_t>(); }
    const uint32_t *run_seg_off() const { return d_run_seg_off.as<uint32_t>(); }
    const uint32_t *scalars() const { return d_scalars.as<uint32_t>(); }
    void release();

    int sort_and_runs(int npass_launch, cudaStream_t st);
};

#ifdef __CUDACC__
__device__ __forceinline__ TileDesc get_tile(const PlanView &P, uint32_t t) {
    if (P.tiles) return P.tiles[t];
    TileDesc d;
    d.seg = 0; d.begin = t * P.tile_elems; d.tile_in_seg = t;
    d.count = (P.N - d.begin < P.tile_elems) ? P.N - d.begin : P.tile_elems;
    return d;
}
__device__ __forceinline__ SegDesc get_seg(const PlanView &P, uint32_t s) {
    if (P.segs) return P.segs[s];
    SegDesc d;
    d.begin = 0; d.count = P.N; d.tile_begin = 0; d.ntiles = P.ntiles;
    return d;
}
#endif

// host cloud (ptr, n, stride, ioff) -> pinned packed float4 staging; returns bbox-free copy
void pack_cloud_f4(const void *src, size_t n, size_t stride, size_t ioff, float *dst_f4);
void pack_cloud_f4_bbox(const void *src, size_t n, size_t stride, size_t ioff, float *dst_f4, float mn[3], float mx[3]);
int key_bits_from_bbox(const float mn[3], const float mx[3], float lx, float ly, float lz);

// largest cloud (or batch of clouds) one call accepts: element positions travel with two flag bits
constexpr size_t B2_MAX_POINTS = 0x3FFFFFF0ull;

}  // namespace b2
