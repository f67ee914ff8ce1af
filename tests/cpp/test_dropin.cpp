// test_dropin.cpp -- exercises the C++ drop-in classes exactly the way the reference's call sites do
// (front_end.cpp:106-107,221-231; matching.cpp:155-158,245-246; loop_closing.cpp:253).
// Usage: test_dropin <target.bin> <scan.bin> <guess16.bin>   (float32 x,y,z,intensity records)
// Prints: M <filtered count> / POSE <16 floats col-major> / FIT <fitness> / ITER <iterations>
#include <cstdio>
#include <fstream>
#include <memory>
#include <vector>

#include "lidar_localization/models/cloud_filter/box_filter.hpp"
#include "lidar_localization/models/cloud_filter/voxel_filter.hpp"
#include "lidar_localization/models/registration/ndt_registration.hpp"

using namespace lidar_localization;

static CloudData::CLOUD_PTR load(const char* path) {
    std::ifstream f(path, std::ios::binary);
    std::vector<float> v((std::istreambuf_iterator<char>(f)), {});
    f.clear(); f.seekg(0, std::ios::end);
    size_t bytes = (size_t)f.tellg(); f.seekg(0);
    v.resize(bytes / 4);
    f.read((char*)v.data(), bytes);
    CloudData::CLOUD_PTR c(new CloudData::CLOUD());
    for (size_t i = 0; i + 3 < v.size(); i += 4) {
        CloudData::POINT p;
        p.x = v[i]; p.y = v[i + 1]; p.z = v[i + 2]; p.intensity = v[i + 3];
        c->push_back(p);
    }
    return c;
}

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage\n"); return 2; }
    CloudData::CLOUD_PTR target = load(argv[1]), scan = load(argv[2]);
    Eigen::Matrix4f guess = Eigen::Matrix4f::Identity();
    { std::ifstream f(argv[3], std::ios::binary); f.read((char*)guess.data(), 64); }

    std::shared_ptr<CloudFilterInterface> frame_filter = std::make_shared<VoxelFilter>(1.3f, 1.3f, 1.3f);
    std::shared_ptr<RegistrationInterface> registration = std::make_shared<NDTRegistration>(1.0f, 0.1f, 0.01f, 30);

    CloudData::CLOUD_PTR filtered(new CloudData::CLOUD());
    frame_filter->Filter(scan, filtered);                 // out pre-allocated empty (front_end.cpp:106)
    CloudData::CLOUD_PTR inplace = load(argv[2]);
    frame_filter->Filter(inplace, inplace);               // in == out (matching.cpp:158)
    if (inplace->points.size() != filtered->points.size()) { std::fprintf(stderr, "in-place mismatch\n"); return 1; }
    for (size_t i = 0; i < filtered->points.size(); ++i)
        if (inplace->points[i].x != filtered->points[i].x || inplace->points[i].intensity != filtered->points[i].intensity) return 1;

    registration->SetInputTarget(target);
    CloudData::CLOUD_PTR result(new CloudData::CLOUD());
    Eigen::Matrix4f pose = Eigen::Matrix4f::Identity();
    registration->ScanMatch(filtered, guess, result, pose);
    float fit = registration->GetFitnessScore();
    std::printf("M %zu\nPOSE", filtered->points.size());
    for (int i = 0; i < 16; ++i) std::printf(" %.9g", pose.data()[i]);
    std::printf("\nFIT %.9g\nITER %d\n", fit, static_cast<NDTRegistration*>(registration.get())->LastResult().iterations);
    // BoxFilter as matching.cpp:166-183 uses it: SetSize, SetOrigin at the pose, Filter, GetEdge
    {
        BoxFilter box(std::vector<float>{-20.f, 20.f, -15.f, 15.f, -2.f, 10.f});
        box.SetOrigin(std::vector<float>{pose(0, 3), pose(1, 3), pose(2, 3)});
        CloudData::CLOUD_PTR cropped(new CloudData::CLOUD());
        box.Filter(target, cropped);
        std::vector<float> e = box.GetEdge();
        size_t expect = 0;
        double sx = 0.0;
        for (const auto& p : target->points)
            if (!(p.x < e[0] || p.y < e[2] || p.z < e[4] || p.x > e[1] || p.y > e[3] || p.z > e[5])) { ++expect; sx += p.x; }
        double gx = 0.0;
        for (const auto& p : cropped->points) gx += p.x;
        std::printf("BOX %zu %zu %.9g %.9g\n", cropped->points.size(), expect, gx, sx);
    }
    // device-resident use of the same classes: crop the map on the device, filter the frame there, match, same answer
    {
        b2cloud *d_map = nullptr, *d_scan = nullptr, *d_filt = nullptr, *d_res = nullptr;
        b2cloud_create(0, &d_map); b2cloud_create(0, &d_scan); b2cloud_create(0, &d_filt); b2cloud_create(0, &d_res);
        b2cloud_upload(d_map, target->points.data(), target->points.size(), 32, 16);
        b2cloud_upload(d_scan, scan->points.data(), scan->points.size(), 32, 16);
        VoxelFilter ff(1.3f, 1.3f, 1.3f);
        ff.FilterDevice(d_scan, d_filt);
        NDTRegistration reg2(1.0f, 0.1f, 0.01f, 30);
        reg2.SetInputTargetDevice(d_map);
        Eigen::Matrix4f pose2 = Eigen::Matrix4f::Identity();
        reg2.ScanMatchDevice(d_filt, guess, d_res, pose2);
        size_t nf = 0, nr = 0;
        b2cloud_size(d_filt, &nf); b2cloud_size(d_res, &nr);
        bool same = nf == filtered->points.size() && nr == nf;
        for (int i = 0; i < 16; ++i) same = same && pose2.data()[i] == pose.data()[i];
        std::printf("DEV %d %d\n", same ? 1 : 0, reg2.LastResult().iterations);
        b2cloud_destroy(d_map); b2cloud_destroy(d_scan); b2cloud_destroy(d_filt); b2cloud_destroy(d_res);
    }
    // UpdateInputTarget (the in-tree NDT's updateVoxelGrid): target set from the first half of the map, second half added;
    // the match must equal the one against the whole map.  The full raw scan as a source exercises the device-side fill
    // of the result cloud (b2ndt_align_ex path of ScanMatch).
    {
        NDTRegistration reg3(1.0f, 0.1f, 0.01f, 30);
        // the points that span the bounding box go into the first part (the update then stays inside the index box)
        CloudData::CLOUD_PTR a(new CloudData::CLOUD()), b(new CloudData::CLOUD());
        size_t ext[6] = {0, 0, 0, 0, 0, 0};
        for (size_t i = 0; i < target->points.size(); ++i) {
            const CloudData::POINT& p = target->points[i];
            if (p.x < target->points[ext[0]].x) ext[0] = i; if (p.x > target->points[ext[1]].x) ext[1] = i;
            if (p.y < target->points[ext[2]].y) ext[2] = i; if (p.y > target->points[ext[3]].y) ext[3] = i;
            if (p.z < target->points[ext[4]].z) ext[4] = i; if (p.z > target->points[ext[5]].z) ext[5] = i;
        }
        for (size_t i = 0; i < target->points.size(); ++i) {
            bool is_ext = false;
            for (int k = 0; k < 6; ++k) is_ext = is_ext || ext[k] == i;
            ((is_ext || i < target->points.size() / 2) ? a : b)->points.push_back(target->points[i]);
        }
        CloudData::CLOUD_PTR whole(new CloudData::CLOUD());
        whole->points = a->points;
        whole->points.insert(whole->points.end(), b->points.begin(), b->points.end());
        bool ok = reg3.SetInputTarget(a) && reg3.UpdateInputTarget(b);
        Eigen::Matrix4f p3 = Eigen::Matrix4f::Identity(), p4 = Eigen::Matrix4f::Identity();
        CloudData::CLOUD_PTR r3(new CloudData::CLOUD()), r4(new CloudData::CLOUD());
        ok = ok && reg3.ScanMatch(scan, guess, r3, p3);
        NDTRegistration reg4(1.0f, 0.1f, 0.01f, 30);
        ok = ok && reg4.SetInputTarget(whole) && reg4.ScanMatch(scan, guess, r4, p4);
        bool same = ok && r3->points.size() == scan->points.size() && r4->points.size() == scan->points.size();
        for (int i = 0; i < 16; ++i) same = same && p3.data()[i] == p4.data()[i];
        // device-filled result cloud == the host formula on the first points
        for (size_t i = 0; same && i < 64 && i < scan->points.size(); ++i) {
            const CloudData::POINT& sp = scan->points[i];
            const float* m = p3.data();
            const float x = ((m[0] * sp.x + m[4] * sp.y) + m[8] * sp.z) + m[12];
            same = same && r3->points[i].x == x && r3->points[i].intensity == sp.intensity && r3->points[i].data[3] == 1.0f;
        }
        std::printf("UPD %d %d %d\n", same ? 1 : 0, reg3.LastResult().iterations, reg4.LastResult().iterations);
    }
    // result cloud = source under the final pose
    if (result->points.size() != filtered->points.size()) return 1;
    std::printf("R0 %.9g %.9g %.9g %.9g\n", result->points[0].x, result->points[0].y, result->points[0].z, result->points[0].intensity);
    return 0;
}
