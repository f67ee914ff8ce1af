// ndt_manual_ref.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" harness around the reference's OWN in-tree NDT
//   /root/reference/lidar_localization/src/models/registration/ndt_registration_manual/
//     {NormalDistributionsTransform,VoxelGrid,Octree,Registration}.cpp
// compiled where they lie (oracle/Makefile target `ref`) against the header stand-ins of
// oracle/ref_stubs/ into oracle/_ref/libndt_manual_ref.so.  Used to pin the oracle's NDT core
// (computeDerivatives / computeAngleDerivatives / computeStepLengthMT / computeTransformation and the
// per-voxel mean / covariance / inverse) against outputs of the reference code itself.
// The private members are reached with the usual test-only access hack.
// everything the reference headers include that is NOT reference code comes first (untouched access)
#include <float.h>
#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include <eigen3/Eigen/Dense>
#include <eigen3/Eigen/Geometry>
#include <boost/make_shared.hpp>
#include <boost/shared_ptr.hpp>
#include <jsk_recognition_msgs/BoundingBox.h>
#include <jsk_recognition_msgs/BoundingBoxArray.h>
#include <pcl/common/common.h>
#include <pcl/common/transforms.h>
#include <pcl/conversions.h>
#include <pcl/filters/extract_indices.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl/kdtree/kdtree_flann.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/search/kdtree.h>
#include <pcl/search/organized.h>
#include <pcl/segmentation/conditional_euclidean_clustering.h>
#include <pcl/segmentation/extract_clusters.h>
#include <pcl/segmentation/sac_segmentation.h>
#include <pcl_conversions/pcl_conversions.h>
#include <pcl_ros/transforms.h>
#include <ros/ros.h>
#include <sensor_msgs/PointCloud2.h>
#include <std_msgs/Header.h>
#define private public
#define protected public
#include "lidar_localization/models/registration/ndt_registration_manual/NormalDistributionsTransform.h"
#undef private
#undef protected

#include <cmath>
#include <cstring>

using lidar_localization::CloudData;
using lidar_localization::NormalDistributionsTransform;

struct Ref {
    NormalDistributionsTransform ndt;
    CloudData::CLOUD_PTR target, source;
};

static CloudData::CLOUD_PTR make_cloud(const float *xyzi, size_t n) {
    CloudData::CLOUD_PTR c(new CloudData::CLOUD());
    c->points.resize(n);
    for (size_t i = 0; i < n; ++i) {
        c->points[i].x = xyzi[4 * i]; c->points[i].y = xyzi[4 * i + 1]; c->points[i].z = xyzi[4 * i + 2];
        c->points[i].intensity = xyzi[4 * i + 3];
    }
    c->width = (uint32_t)n; c->height = 1;
    return c;
}

extern "C" {
void *refndt_new(float res, double step, double eps, int max_iter, double outlier) {
    Ref *r = new Ref();
    r->ndt.setResolution(res);
    r->ndt.setStepSize(step);
    r->ndt.setTransformationEpsilon(eps);
    r->ndt.setMaximumIterations(max_iter);
    r->ndt.setOutlierRatio(outlier);
    return r;
}
void refndt_free(void *h) { delete (Ref *)h; }
void refndt_set_target(void *h, const float *xyzi, size_t n) {
    Ref *r = (Ref *)h;
    r->target = make_cloud(xyzi, n);
    r->ndt.setInputTarget(r->target);
}
// NormalDistributionsTransform::updateVoxelGrid (NormalDistributionsTransform.cpp:968-972 -> VoxelGrid::update,
// VoxelGrid.cpp:545-584, updateVoxelContent :736-809): add a cloud to the current target
void refndt_update(void *h, const float *xyzi, size_t n) {
    Ref *r = (Ref *)h;
    r->ndt.updateVoxelGrid(make_cloud(xyzi, n));
}
// grid geometry: min_b (3), vgrid (3), real min (3), real max (3)
void refndt_grid_info(void *h, int *out12) {
    auto &g = ((Ref *)h)->ndt.voxel_grid_;
    int v[12] = {g.min_b_x_, g.min_b_y_, g.min_b_z_, g.vgrid_x_, g.vgrid_y_, g.vgrid_z_,
                 g.real_min_bx_, g.real_min_by_, g.real_min_bz_, g.real_max_bx_, g.real_max_by_, g.real_max_bz_};
    std::memcpy(out12, v, sizeof(v));
}
// per-voxel data for absolute cell coordinates (ix,iy,iz); returns points_per_voxel
int refndt_voxel(void *h, int ix, int iy, int iz, double *centroid3, double *icov9, double *staticvalue) {
    auto &g = ((Ref *)h)->ndt.voxel_grid_;
    int vid = g.voxelId(ix, iy, iz, g.min_b_x_, g.min_b_y_, g.min_b_z_, g.vgrid_x_, g.vgrid_y_, g.vgrid_z_);
    Eigen::Vector3d c = g.getCentroid(vid);
    Eigen::Matrix3d ic = g.getInverseCovariance(vid);
    for (int a = 0; a < 3; ++a) { centroid3[a] = c(a); for (int b = 0; b < 3; ++b) icov9[a * 3 + b] = ic(a, b); }
    *staticvalue = g.getSaticValue(vid);
    return (*g.points_per_voxel_)[vid];
}
int refndt_radius_search(void *h, float x, float y, float z, float radius, int *ids, int cap) {
    auto &g = ((Ref *)h)->ndt.voxel_grid_;
    CloudData::POINT p; p.x = x; p.y = y; p.z = z;
    std::vector<int> v;
    g.radiusSearch(p, radius, v);
    int n = (int)v.size() < cap ? (int)v.size() : cap;
    for (int i = 0; i < n; ++i) ids[i] = v[i];
    return (int)v.size();
}
// voxel id -> absolute cell coordinates
void refndt_voxel_coords(void *h, int vid, int *ijk) {
    auto &g = ((Ref *)h)->ndt.voxel_grid_;
    int iz = vid / (g.vgrid_x_ * g.vgrid_y_);
    int iy = (vid - iz * g.vgrid_x_ * g.vgrid_y_) / g.vgrid_x_;
    int ix = vid - iz * g.vgrid_x_ * g.vgrid_y_ - iy * g.vgrid_x_;
    ijk[0] = ix + g.min_b_x_; ijk[1] = iy + g.min_b_y_; ijk[2] = iz + g.min_b_z_;
}
void refndt_set_source(void *h, const float *xyzi, size_t n) {
    Ref *r = (Ref *)h;
    r->source = make_cloud(xyzi, n);
    r->ndt.setInputSource(r->source);
}
// computeDerivatives at pose p with the transformed cloud given (n*3 floats); H column-major
double refndt_derivatives(void *h, const float *trans_xyz, const double *p6, int hess, double *g6, double *H36) {
    Ref *r = (Ref *)h;
    // the Gauss constants are normally refreshed by computeTransformation (NDTM:315-321)
    {
        double c1 = 10 * (1 - r->ndt.outlier_ratio_), c2 = r->ndt.outlier_ratio_ / pow(r->ndt.resolution_, 3), d3 = -log(c2);
        r->ndt.gauss_d1_ = -log(c1 + c2) - d3;
        r->ndt.gauss_d2_ = -2 * log((-log(c1 * exp(-0.5) + c2) - d3) / r->ndt.gauss_d1_);
    }
    CloudData::CLOUD trans;
    size_t n = r->source->points.size();
    trans.points.resize(n);
    for (size_t i = 0; i < n; ++i) { trans.points[i].x = trans_xyz[3 * i]; trans.points[i].y = trans_xyz[3 * i + 1]; trans.points[i].z = trans_xyz[3 * i + 2]; }
    Eigen::Matrix<double, 6, 1> g, p;
    Eigen::Matrix<double, 6, 6> H;
    for (int i = 0; i < 6; ++i) p(i) = p6[i];
    double s = r->ndt.computeDerivatives(g, H, trans, p, hess != 0);
    for (int i = 0; i < 6; ++i) { g6[i] = g(i); for (int j = 0; j < 6; ++j) H36[j * 6 + i] = H(i, j); }
    return s;
}
void refndt_angle_tables(void *h, const double *p6, double *j24, double *h45) {
    Ref *r = (Ref *)h;
    Eigen::Matrix<double, 6, 1> p;
    for (int i = 0; i < 6; ++i) p(i) = p6[i];
    r->ndt.computeAngleDerivatives(p, true);
    const Eigen::Vector3d *J[8] = {&r->ndt.j_ang_a_, &r->ndt.j_ang_b_, &r->ndt.j_ang_c_, &r->ndt.j_ang_d_, &r->ndt.j_ang_e_, &r->ndt.j_ang_f_, &r->ndt.j_ang_g_, &r->ndt.j_ang_h_};
    const Eigen::Vector3d *Hh[15] = {&r->ndt.h_ang_a2_, &r->ndt.h_ang_a3_, &r->ndt.h_ang_b2_, &r->ndt.h_ang_b3_, &r->ndt.h_ang_c2_, &r->ndt.h_ang_c3_,
                                     &r->ndt.h_ang_d1_, &r->ndt.h_ang_d2_, &r->ndt.h_ang_d3_, &r->ndt.h_ang_e1_, &r->ndt.h_ang_e2_, &r->ndt.h_ang_e3_,
                                     &r->ndt.h_ang_f1_, &r->ndt.h_ang_f2_, &r->ndt.h_ang_f3_};
    for (int k = 0; k < 8; ++k) for (int a = 0; a < 3; ++a) j24[k * 3 + a] = (*J[k])(a);
    for (int k = 0; k < 15; ++k) for (int a = 0; a < 3; ++a) h45[k * 3 + a] = (*Hh[k])(a);
}
// Registration::align + computeTransformation; pose_out column-major
void refndt_align(void *h, const float *guess16, float *pose16, int *iterations, int *converged, double *trans_prob, float *trans_cloud_xyz) {
    Ref *r = (Ref *)h;
    Eigen::Matrix4f G;
    for (int c = 0; c < 4; ++c) for (int rr = 0; rr < 4; ++rr) G(rr, c) = guess16[c * 4 + rr];
    r->ndt.align(G);
    Eigen::Matrix4f T = r->ndt.getFinalTransformation();
    for (int c = 0; c < 4; ++c) for (int rr = 0; rr < 4; ++rr) pose16[c * 4 + rr] = T(rr, c);
    *iterations = r->ndt.getFinalNumIteration();
    *converged = r->ndt.hasConverged() ? 1 : 0;
    *trans_prob = r->ndt.getTransformationProbability();
    if (trans_cloud_xyz)
        for (size_t i = 0; i < r->ndt.trans_cloud_.points.size(); ++i) {
            trans_cloud_xyz[3 * i] = r->ndt.trans_cloud_.points[i].x; trans_cloud_xyz[3 * i + 1] = r->ndt.trans_cloud_.points[i].y;
            trans_cloud_xyz[3 * i + 2] = r->ndt.trans_cloud_.points[i].z;
        }
}
}
