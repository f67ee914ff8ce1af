"""One single ScanMatch (cluster launch of ndt_match_kernel) + one GetFitnessScore (fitness_kernel) against the 1 M-point
map inside a cudaProfilerStart/Stop range: run under `ncu --profile-from-start off` (launch list / --set full)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lidar_slam_b200 import synth
from lidar_slam_b200.registration import NDTRegistration, VoxelFilter
scene = synth.Scene(leg=500.0)
m = scene.make_map(1_000_000, 2.0)
truth = scene.path_pose(300.0)
scan = scene.scan(4242, truth)
vf = VoxelFilter(1.3, 1.3, 1.3)
_, src = vf.Filter(scan)
guess = synth.pose6_to_matrix(truth + np.array([0.3, -0.2, 0.1, 0.01, -0.01, 0.02])).astype(np.float32)
reg = NDTRegistration(1.0, 0.1, 0.01, 30)
reg.SetInputTarget(m)
for _ in range(3):
    reg.ScanMatch(src, guess); reg.GetFitnessScore()
torch.cuda.synchronize()
torch.cuda.profiler.start()
t = time.perf_counter(); _, _, pose = reg.ScanMatch(src, guess); t_match = 1e3 * (time.perf_counter() - t)
t = time.perf_counter(); fit = reg.GetFitnessScore(); t_fit = 1e3 * (time.perf_counter() - t)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(json.dumps(dict(n_src=len(src), scan_match_ms=t_match, fitness_ms=t_fit, iterations=reg.last_result["iterations"],
                      passes=reg.last_result["passes"], pairs=reg.last_result["pairs"], fitness=fit,
                      err_m=float(np.linalg.norm(pose[:3, 3] - truth[:3])))))
