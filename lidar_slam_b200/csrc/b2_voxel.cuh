// b2_voxel.cuh -- voxelisation pipeline shared by the VoxelFilter and the NDT target-grid build:
//   bbox -> PCL voxel layout -> voxel key per point -> stable LSD radix sort by key (batched over
//   independent clouds) -> run heads (one run = one occupied voxel).
// Replaces pcl::VoxelGrid's std::sort of (idx, point) pairs and VoxelGridCovariance's std::map
// insertion.
#pragma once

#include <vector>

#include "b2_common.cuh"

namespace b2 {

struct VoxPipeline {
    // plan (host + device copies)
    std::vector<TileDesc> h_tiles;
    std::vector<SegDesc> h_segs;
    DevBuf d_tiles, d_segs;
    // per segment
    DevBuf d_bbox;      // uint32[B][8]: ordered-float min xyz, max xyz, n_finite, pad
    DevBuf d_layouts;   // VoxLayout[B]
    DevBuf d_scalars;   // uint32[8]: [0]=max nbits over segments, [1]=total runs
    // per element (ping-pong)
    DevBuf d_keys[2], d_vals[2];
    // radix bookkeeping
    DevBuf d_tilehist;  // uint32[ntiles*256]
    DevBuf d_binbase;   // uint32[B*256]
    // runs
    DevBuf d_tile_heads;   // uint32[ntiles+1]
    DevBuf d_run_start;    // uint32[N+1]   global element index of each run's first element
    DevBuf d_run_seg;      // uint32[N]     cloud of each run
    DevBuf d_run_seg_off;  // uint32[B+1]   first run of each cloud
    size_t N = 0, B = 0, ntiles = 0;
    int final_buf = 0;     // which ping-pong buffer holds the sorted keys/vals

    // Build the plan for B clouds described by host offsets (B+1 entries, in points).
    int plan(const uint32_t *h_offsets, size_t B, cudaStream_t st);
    // Run bbox/layout/keys/sort/run-heads on packed float4 points (device).  nbits_hint > 0 limits the
    // number of radix passes from the host side (must be >= the true key width); 0 = decide on device.
    int run(const float4 *d_pts, float lx, float ly, float lz, int nbits_hint, cudaStream_t st);
    // Generic stable sort + run detection for keys the caller wrote itself: after plan(), fill keys0()/vals0() (one
    // uint32 key per element, key == invalid_key drops the element to the end / out of the runs) and call this with
    // the key width in bits.  Results as after run(): sorted_keys(), sorted_vals(), run_start(), scalars()[1] = runs.
    int run_prepared(uint32_t invalid_key, int nbits, cudaStream_t st);
    uint32_t *keys0() { return d_keys[0].as<uint32_t>(); }
    uint32_t *vals0() { return d_vals[0].as<uint32_t>(); }

    const uint32_t *sorted_keys() const { return d_keys[final_buf].as<uint32_t>(); }
    const uint32_t *sorted_vals() const { return d_vals[final_buf].as<uint32_t>(); }
    const VoxLayout *layouts() const { return d_layouts.as<VoxLayout>(); }
    const uint32_t *run_start() const { return d_run_start.as<uint32_t>(); }
    const uint32_t *run_seg() const { return d_run_seg.as<uint32_t>(); }
    const uint32_t *run_seg_off() const { return d_run_seg_off.as<uint32_t>(); }
    const uint32_t *scalars() const { return d_scalars.as<uint32_t>(); }
    void release();
};

// host cloud (ptr, n, stride, ioff) -> pinned packed float4 staging; returns bbox-free copy
void pack_cloud_f4(const void *src, size_t n, size_t stride, size_t ioff, float *dst_f4);
void pack_cloud_f4_bbox(const void *src, size_t n, size_t stride, size_t ioff, float *dst_f4, float mn[3], float mx[3]);
int key_bits_from_bbox(const float mn[3], const float mx[3], float lx, float ly, float lz);

}  // namespace b2
