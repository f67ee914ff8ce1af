"""VoxelFilter / SetInputTarget on map-sized clouds: wall time through the host API and (under
ncu --profile-from-start off) the per-kernel launch list of one call."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lidar_slam_b200 import synth, capi
from lidar_slam_b200.registration import NDTRegistration, VoxelFilter
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
scene = synth.Scene(leg=500.0)
m = scene.make_map(n, 2.0)
vf = VoxelFilter(0.6, 0.6, 0.6)
reg = NDTRegistration(1.0, 0.1, 0.01, 30)
for _ in range(2):
    vf.Filter(m); reg.SetInputTarget(m)
ts, tt = [], []
for _ in range(5):
    t = time.perf_counter(); ok, f = vf.Filter(m); ts.append(1e3 * (time.perf_counter() - t))
    t = time.perf_counter(); reg.SetInputTarget(m); tt.append(1e3 * (time.perf_counter() - t))
torch.cuda.profiler.start()
vf.Filter(m); reg.SetInputTarget(m)
torch.cuda.profiler.stop()
print(json.dumps({"n": len(m), "n_out": len(f), "filter_ms_p50": float(np.median(ts)), "set_target_ms_p50": float(np.median(tt)),
                  "target": reg.TargetInfo(), "alg_bytes_filter": 16 * len(m) + 16 * len(f)}))
