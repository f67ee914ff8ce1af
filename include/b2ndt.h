/*
 * b2ndt.h -- C ABI of the B200-native NDT scan-matching / voxel-filter engine (libb2ndt.so).
 *
 * This is the drop-in boundary under the reference's plug-in classes.  Each entry point names the
 * reference interface it replaces (paths relative to /root/reference/lidar_localization/):
 *
 *   b2ndt_create / b2ndt_destroy    NDTRegistration ctor + SetRegistrationParam
 *                                   (src/models/registration/ndt_registration.cpp:12-44)
 *   b2ndt_set_target                NDTRegistration::SetInputTarget -> pcl::NDT::setInputTarget
 *                                   (ndt_registration.cpp:46-51; include/.../registration_interface.hpp:18)
 *   b2ndt_align                     NDTRegistration::ScanMatch -> setInputSource + align + getFinalTransformation
 *                                   (ndt_registration.cpp:53-61; registration_interface.hpp:19-22)
 *   b2ndt_fitness                   NDTRegistration::GetFitnessScore -> pcl::Registration::getFitnessScore
 *                                   (ndt_registration.cpp:63-66; registration_interface.hpp:23)
 *   b2ndt_align_batch[_device]      the same ScanMatch for many independent (source, guess) pairs against
 *                                   one target (BASELINE.json configs 4/5; the callers loop ScanMatch per
 *                                   frame: src/matching/matching.cpp:245-246, front_end.cpp:224-231)
 *   b2vf_create / b2vf_destroy      VoxelFilter ctor + SetFilterParam (src/models/cloud_filter/voxel_filter.cpp:12-34)
 *   b2vf_filter                     VoxelFilter::Filter -> pcl::VoxelGrid::filter
 *                                   (voxel_filter.cpp:36-41; include/.../cloud_filter_interface.hpp:17)
 *
 * Conventions: plain pointers and sizes only; return 0 on success, negative b2_status on error
 * (b2_last_error() gives the message for the calling thread); no C++ exceptions cross the ABI.
 * A handle is bound to one CUDA device and one stream and is not thread-safe.  There is no CPU
 * fallback: every entry point fails with B2_ERR_CUDA when no sm_100 device is usable.
 *
 * Host clouds are described by (pointer, count, byte stride, byte offset of intensity) so that
 * pcl::PointXYZI memory (stride 32, xyz at +0, intensity at +16; cloud_data.hpp:35) is consumed in
 * place.  Device clouds are packed float4 {x,y,z,intensity}.  Poses are column-major float[16]
 * (Eigen::Matrix4f memory).
 */
#ifndef B2NDT_H_
#define B2NDT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    B2_OK = 0,
    B2_ERR_INVALID = -1,   /* bad argument */
    B2_ERR_CUDA = -2,      /* CUDA runtime error / no usable device */
    B2_ERR_STATE = -3,     /* call order (e.g. align before set_target) */
    B2_ERR_CAPACITY = -4   /* output buffer too small */
} b2_status;

const char *b2_last_error(void);
/* number of kernels launched by this library in the calling process (all handles) */
uint64_t b2_kernel_launch_count(void);
int b2_device_count(void);

/* ------------------------------------------------------------------ registration ---------- */
typedef struct b2ndt b2ndt;

typedef struct {
    float  res;            /* setResolution           (front_end.yaml:18-22 -> 1.0)  */
    double step_size;      /* setStepSize             (0.1)                          */
    double trans_eps;      /* setTransformationEpsilon(0.01)                         */
    double outlier_ratio;  /* PCL default 0.55                                       */
    int    max_iter;       /* setMaximumIterations    (30)                           */
    int    min_pts;        /* min_points_per_voxel_   (PCL default 6)                */
    double eig_mult;       /* min_covar_eigvalue_mult_(PCL default 0.01)             */
    int    pcl17_compat;   /* 1 (default): More-Thuente interval guard as in PCL 1.7 (loop never runs
                              when step_size > trans_eps/2); 0: full line search enabled */
} b2ndt_params;

typedef struct {
    int32_t iterations;        /* nr_iterations_ */
    int32_t converged;
    double  score;
    double  trans_probability; /* getTransformationProbability() = score / N */
    double  p[6];              /* final (x,y,z,roll,pitch,yaw) */
    int32_t passes;            /* derivative passes executed */
    int32_t mt_trials;
    int64_t pairs;             /* (point, voxel) pairs visited over all passes */
} b2ndt_result;

typedef struct {
    int32_t ok;                /* 0: empty target or PCL's int32 voxel-index guard tripped */
    int32_t min_b[3], div_b[3];
    uint32_t n_points;         /* finite target points */
    uint32_t n_leaves;         /* occupied voxels */
    uint32_t n_tree;           /* voxels with >= min_pts points (searchable) */
    float   inv_leaf;
    uint32_t updates_incremental;  /* b2ndt_update_target calls since the last SetInputTarget that took the incremental path */
    uint32_t updates_rebuilt;      /* ... that had to rebuild (new points outside the grid's index box) */
} b2ndt_target_info;

void b2ndt_params_default(b2ndt_params *p);
int  b2ndt_create(const b2ndt_params *p, int device, b2ndt **out);
void b2ndt_destroy(b2ndt *h);
/* run on an existing CUDA stream (cudaStream_t) instead of the handle's own; NULL restores it */
int  b2ndt_set_stream(b2ndt *h, void *cuda_stream);
int  b2ndt_synchronize(b2ndt *h);
/* thread-block cluster width per match.  single_match_ctas 1..16: upper bound for one ScanMatch (default 16; the
 * library picks by source size).  batch_ctas 0..16: 0 (default) = by batch size: 1 = the persistent batch kernel (two
 * matches in flight per CTA) from ~800 matches on, clusters of 2 / 4 / 8 CTAs per match for smaller batches (a shard of
 * config 4 / 5 on 8 GPUs), as wide as keeps the launch within about two waves of CTAs. */
int  b2ndt_set_cluster(b2ndt *h, int single_match_ctas, int batch_ctas);

int  b2ndt_set_target(b2ndt *h, const void *pts, size_t n, size_t stride, size_t ioff);
int  b2ndt_set_target_device(b2ndt *h, const void *d_pts_f4, size_t n);
/* Incremental target update: NormalDistributionsTransform::updateVoxelGrid(new_cloud) of the reference's in-tree NDT
 * (ndt_registration_manual/NormalDistributionsTransform.cpp:968-972 -> VoxelGrid::update, VoxelGrid.cpp:545-584,
 * updateVoxelContent :736-809): add the points of a new cloud to the target.  The new points are sorted by voxel under
 * the existing layout, every touched voxel continues its stored sums (or a new leaf is opened), only the touched
 * leaves are finished again, and the neighbour lists are laid out again.  The result equals
 * b2ndt_set_target(old points ++ new points) BIT FOR BIT (every per-voxel sum continues in the order the full build
 * would use); when a new point falls outside the grid's index box (PCL's layout of the united cloud differs) the
 * target is rebuilt from all points instead.  The points handed to the last SetInputTarget must still be alive and
 * unchanged at the FIRST update (PCL retains its target pointer the same way); from then on the handle owns a copy.
 * The first b2ndt_fitness* call after an update re-sorts the point buckets (one full build). */
int  b2ndt_update_target(b2ndt *h, const void *pts, size_t n, size_t stride, size_t ioff);
int  b2ndt_update_target_device(b2ndt *h, const void *d_pts_f4, size_t n);
int  b2ndt_target_info_get(b2ndt *h, b2ndt_target_info *info);
/* copy the leaf table to host (ascending voxel index; arrays sized n_leaves, any may be NULL):
 * idx, n_raw, centroid (4 floats), mean (3 doubles), icov (9 doubles, row-major; zeros when n<min_pts) */
int  b2ndt_target_leaves(b2ndt *h, int32_t *idx, int32_t *n_raw, float *centroid4, double *mean3, double *icov9);

int  b2ndt_align(b2ndt *h, const void *src, size_t n, size_t stride, size_t ioff, const float guess[16],
                 float pose_out[16], b2ndt_result *res);
/* b2ndt_align + ScanMatch's result cloud (registration_interface.hpp:19-22; Registration::align(output)): the source
 * transformed by the final pose with pcl::transformPointCloud's float arithmetic, computed on the device (the source is
 * resident there) and written to result_cloud: n records of out_stride bytes, intensity at out_ioff (PointXYZI: 32 / 16,
 * data[3] = 1).  result_cloud may be NULL (= b2ndt_align) or the source buffer. */
int  b2ndt_align_ex(b2ndt *h, const void *src, size_t n, size_t stride, size_t ioff, const float guess[16],
                    float pose_out[16], b2ndt_result *res, void *result_cloud, size_t out_stride, size_t out_ioff);
/* B independent matches against the current target.  Sources are concatenated; offsets has B+1
 * entries (points).  offsets == NULL: one shared source of n_total points x B guesses. */
int  b2ndt_align_batch(b2ndt *h, const void *src, size_t n_total, size_t stride, size_t ioff,
                       const uint32_t *offsets, size_t B, const float *guesses /*B*16*/,
                       float *poses_out /*B*16*/, b2ndt_result *res /*B, may be NULL*/);
/* device-resident variant: packed float4 sources, device offsets (B+1, or NULL), device guesses;
 * results written to device buffers; asynchronous on the handle's stream. */
int  b2ndt_align_batch_device(b2ndt *h, const void *d_src_f4, size_t n_total, const uint32_t *d_offsets,
                              size_t B, const float *d_guesses, float *d_poses_out, b2ndt_result *d_res);
/* one derivative pass at a 6-vector pose (parity gate): H is column-major 6x6 */
int  b2ndt_derivatives(b2ndt *h, const void *src, size_t n, size_t stride, size_t ioff, const double p[6],
                       double *score, double grad[6], double H[36], int64_t *pairs);
/* getFitnessScore of the last b2ndt_align / b2ndt_align_cloud (source + final pose are retained on the device).
 * Any other call that takes a source (align_batch, derivatives, fitness_ex) or a new target ends that state:
 * B2_ERR_STATE until the next single align. */
int  b2ndt_fitness(b2ndt *h, double max_range, double *out);
/* explicit variant: any source / pose against the current target points */
int  b2ndt_fitness_ex(b2ndt *h, const void *src, size_t n, size_t stride, size_t ioff, const float pose[16],
                      double max_range, double *out);

/* ------------------------------------------------------------------ voxel filter ---------- */
typedef struct b2vf b2vf;

int  b2vf_create(float lx, float ly, float lz, int device, b2vf **out);
void b2vf_destroy(b2vf *h);
int  b2vf_set_stream(b2vf *h, void *cuda_stream);
/* out: capacity out_capacity points with byte stride out_stride / intensity at out_ioff (padding bytes
 * of a 32-byte PointXYZI are written as data[3]=1.0f, rest 0).  in == out is allowed (viewer.cpp:207,
 * matching.cpp:158).  *m receives the number of output points.  out_idx / out_cnt (capacity
 * out_capacity, may be NULL) receive the voxel index and the point count of every output point. */
int  b2vf_filter(b2vf *h, const void *in, size_t n, size_t stride, size_t ioff,
                 void *out, size_t out_capacity, size_t out_stride, size_t out_ioff, size_t *m,
                 int32_t *out_idx, int32_t *out_cnt);
/* B clouds in one launch (concatenated, offsets B+1): device float4 in, device float4 out (capacity
 * n_total), d_out_offsets (B+1) receives the output ranges (compacted, same order). Asynchronous. */
int  b2vf_filter_batch_device(b2vf *h, const void *d_in_f4, size_t n_total, const uint32_t *h_offsets,
                              size_t B, void *d_out_f4, uint32_t *d_out_offsets);
/* Batched ingest (raw frames streaming in chunk by chunk: front_end.cpp:106-107 for every frame of a replay): as
 * above, but the filtered clouds are APPENDED to d_out_f4 (capacity out_capacity points) behind what earlier calls
 * left there, with no host round trip: d_cursor[0] (device; zero it before the first call) counts the points in
 * d_out_f4, d_cursor[1] is set when a batch did not fit (that batch is dropped as a whole).
 * d_out_offsets[first_index + s], s = 0..B, receives where cloud s of this batch starts: after K calls the array is
 * the offsets table b2ndt_align_batch_device takes.  Asynchronous. */
int  b2vf_filter_batch_append_device(b2vf *h, const void *d_in_f4, size_t n_total, const uint32_t *h_offsets,
                                     size_t B, void *d_out_f4, size_t out_capacity, uint32_t *d_out_offsets,
                                     size_t first_index, uint32_t *d_cursor);

/* ------------------------------------------------------------------ device-resident clouds --
 * The callers either side of the hot path keep their clouds in HBM: the front end's
 * local map = sum of key frames transformed by their poses (front_end.cpp:375-410), the matching node's
 * BoxFilter crop of the global map (box_filter.cpp:27-37, matching.cpp:166-183), VoxelFilter and
 * SetInputTarget / ScanMatch on the result -- no host round trip between the steps.  A b2cloud is a packed
 * float4 {x,y,z,intensity} array on one device; handles are not thread-safe. */
typedef struct b2cloud b2cloud;

int  b2cloud_create(int device, b2cloud **out);
void b2cloud_destroy(b2cloud *c);
int  b2cloud_upload(b2cloud *c, const void *pts, size_t n, size_t stride, size_t ioff);
/* out: capacity points with byte stride / intensity offset as in b2vf_filter; *n receives the size */
int  b2cloud_download(b2cloud *c, void *out, size_t capacity, size_t stride, size_t ioff, size_t *n);
int  b2cloud_size(b2cloud *c, size_t *n);
int  b2cloud_clear(b2cloud *c);
int  b2cloud_device_ptr(b2cloud *c, void **d_f4);
/* dst += pcl::transformPointCloud(src, T) (T column-major float 4x4; intensity kept; order kept) */
int  b2cloud_append_transformed(b2cloud *dst, b2cloud *src, const float T[16]);
/* the whole local-map assembly of front_end.cpp:398-407 in one launch: dst = srcs[0] moved by poses[0] ++ srcs[1] moved
 * by poses[1] ++ ... (K column-major float[16] poses); same points, same order as K calls of the function above */
int  b2cloud_assemble(b2cloud *dst, b2cloud *const *srcs, const float *poses, size_t K);
/* pcl::CropBox: dst = points of src with edge[0] <= x <= edge[1], edge[2] <= y <= edge[3],
 * edge[4] <= z <= edge[5] (BoxFilter::GetEdge order), input order kept, non-finite points dropped.
 * For this call, b2cloud_remove_nan and b2cloud_distortion_adjust dst may be src (the reference's in == out calls):
 * the survivors are written to the cloud's second buffer and the buffers swapped, so b2cloud_device_ptr changes. */
int  b2cloud_box_filter(b2cloud *src, const float edge[6], b2cloud *dst);
/* pcl::removeNaNFromPointCloud: dst = points of src with finite x, y, z, input order kept (front_end.cpp:92) */
int  b2cloud_remove_nan(b2cloud *src, b2cloud *dst);
/* DistortionAdjust::SetMotionInfo + AdjustCloud (src/models/scan_adjust/distortion_adjust.cpp:10-69): undo the sensor's
 * motion inside one sweep of scan_period seconds (velocities in the sensor frame); drops point 0 and the 5 degree
 * sector around the first point's azimuth, output intensity is 0, input order kept */
int  b2cloud_distortion_adjust(b2cloud *src, float scan_period, const double linear_velocity[3],
                               const double angular_velocity[3], b2cloud *dst);
/* VoxelFilter::Filter on device clouds (src == dst allowed) */
int  b2vf_filter_cloud(b2vf *h, b2cloud *src, b2cloud *dst);
/* Fused ingest of a raw scan in HBM: DistortionAdjust::AdjustCloud (distortion_adjust.cpp:16-69; both velocities NULL
 * = no de-skew) + pcl::removeNaNFromPointCloud (front_end.cpp:92) + VoxelFilter::Filter (front_end.cpp:106-107) with
 * the first two folded into the first kernel of the filter (one pass over the raw points, no compaction, no extra
 * host round trip).  filtered == VoxelFilter(removeNaN(deskew(src))) bit for bit.  ingested (may be NULL; not src)
 * receives the de-skewed scan with src's size and NaN points where a point was dropped: consumers of device clouds
 * skip non-finite points, b2cloud_remove_nan compacts them away. */
int  b2vf_ingest_filter_cloud(b2vf *h, b2cloud *src, float scan_period, const double linear_velocity[3],
                              const double angular_velocity[3], b2cloud *filtered, b2cloud *ingested);
/* SetInputTarget / ScanMatch on device clouds; result_cloud may be NULL */
int  b2ndt_set_target_cloud(b2ndt *h, b2cloud *target);
int  b2ndt_update_target_cloud(b2ndt *h, b2cloud *add);     /* see b2ndt_update_target */
int  b2ndt_align_cloud(b2ndt *h, b2cloud *src, const float guess[16], float pose_out[16], b2ndt_result *res,
                       b2cloud *result_cloud);

/* ------------------------------------------------------------------ PCD files ----------------
 * PCD v0.7 I/O for PointXYZI clouds (pcl::io::loadPCDFile at matching.cpp:155, loop_closing.cpp:134,286,304;
 * pcl::io::savePCDFileBinary at back_end.cpp:194, viewer.cpp:202,210).  DATA ascii and binary are read (any field
 * set containing x y z, intensity optional), binary PointXYZI is written; binary_compressed (LZF) files are read as well.
 * b2_pcd_read returns a malloc'ed packed {x,y,z,intensity} array (release with b2_pcd_free); host-only calls. */
int  b2_pcd_read(const char *path, float **xyzi, size_t *n_points);
void b2_pcd_free(float *xyzi);
int  b2_pcd_write_binary(const char *path, const float *xyzi, size_t n_points);
int  b2cloud_load_pcd(b2cloud *c, const char *path);
int  b2cloud_save_pcd(b2cloud *c, const char *path);

/* ------------------------------------------------------------------ initial-yaw search -------
 * The matching node's position-only initialisation: Matching::generateGauss2DMapCells
 * (matching.cpp:344-394: 2-D grid over the local map minus its origin, per cell the running mean / variance of z
 * in input order) and Matching::getInitialYawAngle (matching.cpp:267-308: angle_size yaw bins, the scan rotated
 * about z and scored against the grid, first maximal bin wins).  grid_resolution is the reference's double
 * `grid_resolution` (clamped to >= 0.1, matching.cpp:138). */
typedef struct b2hmap b2hmap;

int  b2hmap_create(int device, double grid_resolution, b2hmap **out);
void b2hmap_destroy(b2hmap *h);
int  b2hmap_build(b2hmap *h, b2cloud *local_map, const float origin[3]);
int  b2hmap_info(b2hmap *h, int32_t *width, int32_t *height, float min_xyz[3], float max_xyz[3]);
/* cell arrays [width][height] (cell = x * height + y): mu, sigma, point_cnt; any may be NULL */
int  b2hmap_cells(b2hmap *h, float *mu, float *sigma, int32_t *point_cnt);
/* probs (angle_size doubles, may be NULL) receives the per-bin scores, *best_angle the winning yaw in radians */
int  b2hmap_yaw_search(b2hmap *h, b2cloud *scan, int angle_size, double *probs, double *best_angle);
/* the same score for (x, y, yaw) hypotheses: n_offsets sensor positions (dx, dy relative to the grid origin) x
 * angle_size yaw bins -> probs[o * angle_size + bin] (BASELINE.json config 5: hypotheses for batched NDT) */
int  b2hmap_pose_search(b2hmap *h, b2cloud *scan, int angle_size, const float *offsets_xy, int n_offsets, double *probs);

#ifdef __cplusplus
}
#endif
#endif
