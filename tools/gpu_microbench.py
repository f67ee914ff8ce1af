"""Per-entry-point timings on one B200 (host-buffer C ABI, wall clock incl. copies), for BASELINE.md / profiles."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lidar_slam_b200 import synth, capi
from lidar_slam_b200.registration import NDTRegistration, VoxelFilter


def med(f, n=15, warm=3):
    for _ in range(warm):
        f()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); f(); ts.append(1e3 * (time.perf_counter() - t))
    return float(np.median(ts)), float(np.min(ts))


scene = synth.Scene(leg=500.0)
out = {}
m1 = scene.make_map(1_000_000, 2.0)
p = scene.path_pose(300.0)
scan = scene.scan(4242, p)
for leaf in (1.3, 0.6):
    vf = VoxelFilter(leaf, leaf, leaf)
    l0 = capi.launches(); ok, f = vf.Filter(scan); nl = capi.launches() - l0
    out["voxel_filter_scan_%dk_leaf%.1f" % (len(scan) // 1000, leaf)] = dict(zip(("p50_ms", "min_ms"), med(lambda: vf.Filter(scan))), n_out=len(f), launches=nl)
vf = VoxelFilter(0.6, 0.6, 0.6)
out["voxel_filter_map_1M_leaf0.6"] = dict(zip(("p50_ms", "min_ms"), med(lambda: vf.Filter(m1), 7, 2)), n_out=len(vf.Filter(m1)[1]))
reg = NDTRegistration(1.0, 0.1, 0.01, 30)
out["set_target_1M"] = dict(zip(("p50_ms", "min_ms"), med(lambda: reg.SetInputTarget(m1), 9, 2)), info=reg.TargetInfo())
src = VoxelFilter(1.3, 1.3, 1.3).Filter(scan)[1]
rng = np.random.default_rng(3)
guess = synth.pose6_to_matrix(synth.perturb_pose(p, rng)).astype(np.float32)
for C_ in (1, 2, 4, 8, 10, 11, 12, 16):
    reg.SetCluster(C_, 1)
    t = med(lambda: reg.ScanMatch(src, guess, want_cloud=False), 31, 5)
    out["align_cluster%d" % C_] = dict(p50_ms=t[0], min_ms=t[1], iterations=reg.last_result["iterations"], n_src=len(src))
reg.SetCluster(16, 1)
out["fitness"] = dict(zip(("p50_ms", "min_ms"), med(lambda: reg.GetFitnessScore())))
out["align_raw_scan_120k_cluster16"] = None
reg.SetCluster(16, 1)
t = med(lambda: reg.ScanMatch(scan, guess, want_cloud=False), 7, 2)
out["align_raw_scan_120k_cluster16"] = dict(p50_ms=t[0], min_ms=t[1], iterations=reg.last_result["iterations"], n_src=len(scan))
# position-only initialisation of the matching node (matching.cpp:327-342): crop +-100 m, height grid, 270-bin yaw scan
from lidar_slam_b200.registration import BoxFilter, DeviceCloud, InitialYawSearch
d_map = DeviceCloud(m1)
box = BoxFilter([-100.0, 100.0, -100.0, 100.0, -100.0, 100.0]); box.SetOrigin(p[:3])
d_local = box.FilterCloud(d_map)
ys = InitialYawSearch(0.8)
out["height_grid_build"] = dict(zip(("p50_ms", "min_ms"), med(lambda: ys.GenerateGauss2DMapCells(d_local, p[:3]), 9, 2)),
                                local_map_points=len(d_local), info={k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in ys.Info().items()})
d_scan = DeviceCloud(scan)
out["yaw_search_270_bins"] = dict(zip(("p50_ms", "min_ms"), med(lambda: ys.GetInitialYawAngle(d_scan, 270), 9, 2)), n_scan=len(scan),
                                  best_yaw=ys.GetInitialYawAngle(d_scan, 270)[0], truth_yaw=float(p[5]))
print(json.dumps(out, indent=1))
