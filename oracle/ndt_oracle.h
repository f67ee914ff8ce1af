/*
 * ndt_oracle.h -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the arithmetic behind the reference's registration hot path
 * (lidar_localization NDTRegistration / VoxelFilter, which forward to PCL 1.7):
 *   pcl::VoxelGrid<PointXYZI>::applyFilter           (voxel_filter.cpp:36-41 call site)
 *   pcl::VoxelGridCovariance<PointXYZI>::applyFilter (ndt_registration.cpp:46-51 call site)
 *   pcl::NormalDistributionsTransform::computeTransformation & friends
 *                                                    (ndt_registration.cpp:53-61 call site)
 *   pcl::Registration::getFitnessScore               (ndt_registration.cpp:63-66 call site)
 * PCL 1.7.2 itself is NOT vendored in /root/reference and not installed; the algorithm text
 * followed here is the in-tree hand-written copy
 *   lidar_localization/src/models/registration/ndt_registration_manual/NormalDistributionsTransform.cpp
 *   lidar_localization/src/models/registration/ndt_registration_manual/VoxelGrid.cpp
 * with the PCL-1.7-vs-in-tree deltas of SURVEY.md Appendix A.4 applied, and Eigen 3.2.92
 * numerics (lidar_localization/third_party/eigen3) restated in C.
 *
 * PARITY STATUS: the reference ships no tests / golden vectors for this path,
 * so the oracle is pinned only (a) on its Eigen numerics against the real vendored Eigen
 * (oracle/_ref/libeigen_ref.so, built by oracle/Makefile from the headers where they lie) and
 * (b) on the in-tree NDT source compiled against stub headers (oracle/_ref/libndt_manual_ref.so,
 * see oracle/Makefile) in its "in-tree compat" mode.  The PCL-specific deltas are unpinned.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * use anything in this directory.  The product (lidar_slam_b200/) never links or imports it.
 */
#ifndef NDT_ORACLE_H_
#define NDT_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* A cloud is described by (base pointer, count, byte stride, byte offset of intensity).
 * pcl::PointXYZI: stride 32, xyz at +0, intensity at +16 (cloud_data.hpp:35).
 * packed float4 {x,y,z,i}: stride 16, intensity at +12. */
typedef struct {
    const void *data;
    size_t      n;
    size_t      stride;
    size_t      ioff;
} orc_cloud;

/* ---- voxel index layout shared by VoxelGrid and VoxelGridCovariance (Appendix A.2/A.3) ---- */
typedef struct {
    int     ok;          /* 0: PCL's "leaf size too small" int32 guard tripped (or no finite point) */
    float   inv[3];      /* inverse_leaf_size_ = 1.0f / leaf */
    float   min_p[3], max_p[3];
    int32_t min_b[3], max_b[3], div_b[3], divb_mul[3];
    size_t  n_finite;
} orc_vox_layout;

int  orc_vox_layout_compute(orc_cloud c, float lx, float ly, float lz, orc_vox_layout *out);
/* voxel linear index of one point under a layout (int32, PCL formula). */
int32_t orc_vox_index(const orc_vox_layout *L, float x, float y, float z);

/* ---- pcl::VoxelGrid (down-sampling filter) ----
 * out_xyzi: capacity n*4 floats; out_idx/out_cnt: capacity n (may be NULL).
 * Returns number of output points M (ascending voxel index).  On the overflow guard PCL copies
 * the input to the output: returns n with *overflow=1 and out = input points in input order.
 * Within-voxel accumulation order is the input order (PCL's std::sort is unstable, so any order
 * is a valid PCL outcome; see Appendix A.3). */
size_t orc_voxel_filter(orc_cloud in, float lx, float ly, float lz,
                        float *out_xyzi, int32_t *out_idx, int32_t *out_cnt, int *overflow);

/* ---- pcl::VoxelGridCovariance (NDT target cells) ---- */
typedef struct {
    int32_t idx;        /* linear voxel index */
    int32_t n_raw;      /* points that fell into the voxel */
    int32_t nr_points;  /* PCL's leaf.nr_points after the eigen checks: n_raw, or -1 */
    int32_t in_tree;    /* 1 when n_raw >= min_points_per_voxel: part of the centroid kd-tree */
    float   centroid[4];/* float-accumulated x,y,z,intensity mean (kd-tree search point) */
    double  mean[3];
    double  cov[9];     /* row-major 3x3 (after inflation) */
    double  icov[9];
    double  evals[3];
} orc_leaf;

typedef struct orc_grid orc_grid;

orc_grid *orc_grid_build(orc_cloud target, float resolution, int min_pts, double eig_mult);
void      orc_grid_free(orc_grid *g);
size_t    orc_grid_num_leaves(const orc_grid *g);   /* all occupied voxels, ascending idx */
const orc_leaf *orc_grid_leaves(const orc_grid *g);
const orc_vox_layout *orc_grid_layout(const orc_grid *g);
/* radiusSearch on the float centroid cloud: strict d2 < (float)(r*r), float L2; results sorted by
 * (d2, leaf slot).  Returns count; fills up to cap slots (indices into orc_grid_leaves). */
int orc_grid_radius_search(const orc_grid *g, float qx, float qy, float qz, double radius,
                           int32_t *slots, float *d2, int cap);

/* ---- numerics (Eigen 3.2.92 restated) ---- */
void orc_euler_angles_012_f32(const float T_colmajor4x4[16], float out[3]);
void orc_pose_to_matrix_f32(const double p[6], float T_colmajor4x4[16]);
/* 1: float sin/cos through libm sinf/cosf (what the reference build calls); 0 (default): through
 * float(sin(double)) which both this oracle and the CUDA path can reproduce bit for bit. */
void orc_set_f32_trig_libm(int on);
/* JacobiSVD<Matrix<double,6,6>>(H, FullU|FullV).solve(b); H column-major. Returns rank. */
int  orc_jacobi_svd_solve6(const double H[36], const double b[6], double x[6], double sv[6]);
/* symmetric 3x3 eigen decomposition, ascending eigenvalues; A and evecs row-major (evecs columns
 * are eigenvectors). */
void orc_eig3_sym(const double A[9], double evals[3], double evecs[9]);
void orc_inverse3(const double A[9], double out[9]);
void orc_gauss_constants(double outlier_ratio, float resolution, double *d1, double *d2);
/* pcl::transformPointCloud with a float Matrix4f: one point. */
void orc_transform_point_f32(const float T[16], float x, float y, float z, float out[3]);

/* ---- NDT ---- */
typedef struct {
    float  res;
    double step_size;
    double trans_eps;
    double outlier_ratio;   /* PCL default 0.55 */
    int    max_iter;
    int    min_pts;         /* 6 */
    double eig_mult;        /* 0.01 */
    int    pcl17_compat;    /* 1: More-Thuente loop guard as written in PCL 1.7
                               (interval_converged = (step_max-step_min) > 0 => loop dead);
                               0: loop enabled (interval_converged starts false). */
} orc_params;

void orc_params_default(orc_params *p);

/* computeDerivatives at pose p (6-vector) given the already transformed source (trans_xyz:
 * n*3 floats).  H is column-major 6x6 (symmetric up to rounding).  Returns score.
 * pairs (may be NULL) receives the number of (point, voxel) pairs visited. */
double orc_ndt_derivatives(const orc_grid *g, const orc_params *prm, orc_cloud src,
                           const float *trans_xyz, const double p[6], int compute_hessian,
                           double grad[6], double H[36], long long *pairs);

typedef struct {
    int    iterations;      /* nr_iterations_ */
    int    converged;
    double score;
    double trans_probability;
    double p[6];            /* final 6-vector */
    int    passes;          /* derivative passes executed */
    long long pairs;        /* (point,voxel) pairs over all passes */
    int    mt_trials;       /* More-Thuente trial evaluations (0 in pcl17_compat) */
} orc_result;

/* Registration::align + NDT::computeTransformation.  guess/pose_out are column-major float 4x4.
 * result_xyz (n*3 floats, may be NULL) receives the source transformed by the final pose (the
 * `output` cloud of align).  trace (may be NULL, capacity trace_cap rows of 8 doubles):
 * per Newton iteration {p0..p5, score, step}. */
int orc_ndt_align(const orc_grid *g, const orc_params *prm, orc_cloud src, const float guess[16],
                  float pose_out[16], float *result_xyz, orc_result *res,
                  double *trace, int trace_cap);

/* Registration::getFitnessScore(max_range): mean float squared distance from each transformed
 * source point to its exact nearest target point (all finite target points). */
double orc_fitness_score(orc_cloud target, orc_cloud src, const float pose[16], double max_range);

#ifdef __cplusplus
}
#endif
#endif
