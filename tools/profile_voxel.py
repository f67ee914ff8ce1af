"""One device-resident VoxelFilter::Filter (raw scan, 1 M and 5 M-point maps) and one SetInputTarget (1 M, 5 M) inside a
cudaProfilerStart/Stop range: run under `ncu --profile-from-start off` for the per-kernel launch list / --set full."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lidar_slam_b200 import synth
from lidar_slam_b200.registration import NDTRegistration, VoxelFilter, DeviceCloud
which = sys.argv[1] if len(sys.argv) > 1 else "all"
scene = synth.Scene(leg=500.0)
scan = scene.scan(4242, scene.path_pose(300.0))
clouds = {"scan": (DeviceCloud(scan), 1.3)}
if which in ("all", "1m"):
    clouds["1m"] = (DeviceCloud(scene.make_map(1_000_000, 2.0)), 0.6)
if which in ("all", "5m"):
    clouds["5m"] = (DeviceCloud(scene.make_map(5_000_000, 2.0)), 0.6)
reg = NDTRegistration(1.0, 0.1, 0.01, 30)
d_f = DeviceCloud()
vfs = {k: VoxelFilter(leaf, leaf, leaf) for k, (c, leaf) in clouds.items()}
for _ in range(3):
    for k, (c, leaf) in clouds.items():
        vfs[k].FilterCloud(c, d_f)
        if k != "scan":
            reg.SetInputTargetCloud(c)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = {}
for k, (c, leaf) in clouds.items():
    t = time.perf_counter(); vfs[k].FilterCloud(c, d_f); out["filter_" + k + "_ms"] = 1e3 * (time.perf_counter() - t)
    out["filter_" + k + "_n"] = [len(c), len(d_f)]
    if k != "scan":
        t = time.perf_counter(); reg.SetInputTargetCloud(c); out["target_" + k + "_ms"] = 1e3 * (time.perf_counter() - t)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(json.dumps(out))
