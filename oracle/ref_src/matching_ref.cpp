// matching_ref.cpp -- TEST INFRASTRUCTURE ONLY.  C entry points around the reference's OWN position-only initialisation
// (lidar_localization/src/matching/matching.cpp: generateGauss2DMapCells :344-394, getInitialYawAngle :267-308,
// resetMapRange), compiled where it lies by oracle/Makefile against the header stand-ins of oracle/ref_stubs and the
// vendored Eigen 3.2.92: pins the oracle's restatement of the height grid and of the 270-bin yaw scan.
// Matching's constructor runs against a prepared YAML tree (registration "NDT", no_filter everywhere) and stand-in
// registration / filter classes; the map file does not exist, so its maps stay empty.
#include <cstring>
#include <limits>
#include <yaml-cpp/yaml.h>
#define private public
#include "lidar_localization/matching/matching.hpp"
#undef private

using namespace lidar_localization;

static CloudData::CLOUD_PTR make_cloud(const float *xyzi, size_t n) {
    CloudData::CLOUD_PTR c(new CloudData::CLOUD());
    c->points.resize(n);
    for (size_t i = 0; i < n; ++i) {
        c->points[i].x = xyzi[4 * i]; c->points[i].y = xyzi[4 * i + 1]; c->points[i].z = xyzi[4 * i + 2];
        c->points[i].intensity = xyzi[4 * i + 3];
    }
    c->width = (unsigned)n; c->height = 1;
    return c;
}

extern "C" {
void *refmatch_new(double grid_resolution) {
    YAML::Node &t = YAML::b2_loadfile_tree();
    t = YAML::Node();
    t.set("map_path", YAML::Node(std::string("/nonexistent/map.pcd")));
    t.set("registration_method", YAML::Node(std::string("NDT")));
    t.set("NDT", YAML::Node(std::string("")));
    t.set("global_map_filter", YAML::Node(std::string("no_filter")));
    t.set("local_map_filter", YAML::Node(std::string("no_filter")));
    t.set("frame_filter", YAML::Node(std::string("no_filter")));
    t.set("init_type", YAML::Node(std::string("OnlyPosition")));
    t.set("grid_resolution", YAML::Node(grid_resolution));
    return new Matching();
}
void refmatch_free(void *h) { delete (Matching *)h; }
// what SetInitPose does around generateGauss2DMapCells (matching.cpp:327-342), with the local map given directly
void refmatch_build(void *h, const float *xyzi, size_t n, const float origin[3]) {
    Matching *m = (Matching *)h;
    m->local_map_ptr_ = make_cloud(xyzi, n);
    m->local_map_origion_ = Eigen::Vector3f(origin[0], origin[1], origin[2]);
    m->map_min_xyz_ = Eigen::Vector3f(std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max());
    m->map_max_xyz_ = Eigen::Vector3f(-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max());
    m->generateGauss2DMapCells();
}
void refmatch_info(void *h, int *wh, float *min3, float *max3) {
    Matching *m = (Matching *)h;
    wh[0] = m->local_map_width_; wh[1] = m->local_map_height_;
    for (int a = 0; a < 3; ++a) { min3[a] = m->map_min_xyz_(a); max3[a] = m->map_max_xyz_(a); }
}
// cells in [x][y] order, x-major
void refmatch_cells(void *h, float *mu, float *sigma, int *cnt) {
    Matching *m = (Matching *)h;
    size_t k = 0;
    for (int x = 0; x < m->local_map_width_; ++x)
        for (int y = 0; y < m->local_map_height_; ++y, ++k) {
            mu[k] = m->map_cell_datas_[x][y].mu; sigma[k] = m->map_cell_datas_[x][y].sigma; cnt[k] = m->map_cell_datas_[x][y].point_cnt;
        }
}
double refmatch_yaw(void *h, const float *xyzi, size_t n) {
    Matching *m = (Matching *)h;
    return m->getInitialYawAngle(make_cloud(xyzi, n));
}
}
