"""Device-resident VoxelFilter / SetInputTarget timings (wall clock per call incl. the call's own sync) on one B200:
raw scan, 1 M and 5 M point maps; prints JSON.  Also checks ids / counts / centroids against the oracle on the scan."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lidar_slam_b200 import synth, capi
from lidar_slam_b200.registration import NDTRegistration, VoxelFilter, DeviceCloud


def med(f, n=15, warm=3):
    for _ in range(warm):
        f()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); f(); ts.append(1e3 * (time.perf_counter() - t))
    return float(np.median(ts)), float(np.min(ts))


PEAK = 6534.8
scene = synth.Scene(leg=500.0)
out = {}
p = scene.path_pose(300.0)
scan = scene.scan(4242, p)
sizes = [int(a) for a in sys.argv[1:]] or [1_000_000, 5_000_000]
d_scan, d_f = DeviceCloud(scan), DeviceCloud()
for leaf in (1.3, 0.6):
    vf = VoxelFilter(leaf, leaf, leaf)
    l0 = capi.launches(); vf.FilterCloud(d_scan, d_f); nl = capi.launches() - l0
    t = med(lambda: vf.FilterCloud(d_scan, d_f), 31, 5)
    b = 16.0 * (len(scan) + len(d_f))
    out["filter_scan_%dk_leaf%.1f" % (len(scan) // 1000, leaf)] = dict(p50_ms=t[0], min_ms=t[1], n_out=len(d_f), launches=nl,
                                                                       GBps=b / t[0] / 1e6, frac=b / t[0] / 1e6 / PEAK)
reg = NDTRegistration(1.0, 0.1, 0.01, 30)
t = med(lambda: reg.SetInputTargetCloud(d_scan), 15, 3)
out["set_target_scan"] = dict(p50_ms=t[0], min_ms=t[1], info=reg.TargetInfo())
for n in sizes:
    m = scene.make_map(n, 2.0)
    d_m = DeviceCloud(m)
    vf = VoxelFilter(0.6, 0.6, 0.6)
    t = med(lambda: vf.FilterCloud(d_m, d_f), 15, 3)
    b = 16.0 * (len(m) + len(d_f))
    out["filter_map_%dM_leaf0.6" % (n // 1_000_000)] = dict(p50_ms=t[0], min_ms=t[1], n_out=len(d_f), GBps=b / t[0] / 1e6, frac=b / t[0] / 1e6 / PEAK)
    l0 = capi.launches(); reg.SetInputTargetCloud(d_m); nl = capi.launches() - l0
    t = med(lambda: reg.SetInputTargetCloud(d_m), 15, 3)
    info = reg.TargetInfo()
    b = 16.0 * len(m) + 80.0 * info["n_leaves"]
    out["set_target_%dM" % (n // 1_000_000)] = dict(p50_ms=t[0], min_ms=t[1], launches=nl, GBps=b / t[0] / 1e6, frac=b / t[0] / 1e6 / PEAK, info=info)
    if n == sizes[0]:
        # incremental update: a raw scan taken inside the map added to the target (b2ndt_update_target_cloud), against the
        # full build of map ++ scan; each timed update starts from a freshly built target
        box = reg.TargetInfo()
        ijk = np.floor(scan[:, :3] * np.float32(1.0)).astype(np.int64) - np.array(box["min_b"])
        scan_in = np.ascontiguousarray(scan[((ijk >= 0) & (ijk < np.array(box["div_b"]))).all(1)])   # the part inside the grid's index box
        d_scan_in = DeviceCloud(scan_in)
        d_both = DeviceCloud(np.concatenate([m, scan_in]))
        t_full = med(lambda: reg.SetInputTargetCloud(d_both), 9, 2)
        half = len(scan_in) // 2
        d_h1, d_h2 = DeviceCloud(scan_in[:half]), DeviceCloud(scan_in[half:])
        ts, ts1, ts2, path = [], [], [], None
        for _ in range(9):
            reg.SetInputTargetCloud(d_m)
            t0 = time.perf_counter(); reg.UpdateInputTarget(d_scan_in); ts.append(1e3 * (time.perf_counter() - t0))
            path = reg.TargetInfo()
            # first update after a SetInputTarget (takes its copy of the target's points) / every later one
            reg.SetInputTargetCloud(d_m)
            t0 = time.perf_counter(); reg.UpdateInputTarget(d_h1); ts1.append(1e3 * (time.perf_counter() - t0))
            t0 = time.perf_counter(); reg.UpdateInputTarget(d_h2); ts2.append(1e3 * (time.perf_counter() - t0))
        out["update_target_%dM_plus_scan" % (n // 1_000_000)] = dict(p50_ms=float(np.median(ts[2:])), min_ms=float(np.min(ts)), full_build_ms=t_full[0],
                                                                    first_half_ms=float(np.median(ts1[2:])), second_half_ms=float(np.median(ts2[2:])),
                                                                    incremental=path["updates_incremental"], rebuilt=path["updates_rebuilt"],
                                                                    n_leaves=path["n_leaves"], n_leaves_before=box["n_leaves"], n_added=len(scan_in))
        del d_both
    del d_m
print(json.dumps(out, indent=1, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o)))
