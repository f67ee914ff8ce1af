"""Verbose first-contact check on a GPU box: voxel filter, target grid, derivatives, align, fitness vs oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lidar_slam_b200 import synth
from lidar_slam_b200.registration import NDTRegistration, VoxelFilter
from oracle import oracle as O

scene = synth.Scene(leg=100.0)
m = scene.make_map(300000, 2.0)
p = scene.path_pose(41.3)
s = scene.scan(1000, p)
t = time.time(); vf = VoxelFilter(1.3, 1.3, 1.3); print("vf create", time.time() - t)
for rep in range(3):
    t = time.time(); ok, f, idx, cnt = vf.Filter(s, with_info=True); print("filter", time.time() - t, f.shape)
of, oidx, ocnt, _ = O.voxel_filter(s, 1.3, 1.3, 1.3)
print("filter idx eq", np.array_equal(idx, oidx), "cnt eq", np.array_equal(cnt, ocnt), "centroid eq", np.array_equal(f, of),
      "max diff", np.abs(f - of).max() if f.shape == of.shape else None)
reg = NDTRegistration(1.0, 0.1, 0.01, 30)
for rep in range(2):
    t = time.time(); reg.SetInputTarget(m); print("set target", time.time() - t)
print(reg.TargetInfo())
g = O.Grid(m, 1.0); lv = g.leaves(); L = reg.TargetLeaves()
print("leaves", len(lv), len(L["idx"]), "idx eq", np.array_equal(lv["idx"], L["idx"]), "n eq", np.array_equal(lv["n_raw"], L["n"]),
      "centroid eq", np.array_equal(lv["centroid"], L["centroid"]), "mean eq", np.array_equal(lv["mean"], L["mean"]))
tree = lv["n_raw"] >= 6
d = np.abs(lv["icov"][tree] - L["icov"][tree]); sc_ = np.abs(lv["icov"][tree]).max(1, keepdims=True)
print("icov bit-equal frac", np.mean(np.all(lv["icov"][tree] == L["icov"][tree], axis=1)), "max rel", (d / sc_).max())
prm = O.params(step_size=float(np.float32(0.1)), trans_eps=float(np.float32(0.01)))
rng = np.random.default_rng(1)
for trial in range(4):
    gp = synth.perturb_pose(p, rng)
    sc0, g0, H0, pr0 = O.derivatives(g, prm, of, gp)
    sc1, g1, H1, pr1 = reg.Derivatives(of, gp)
    print("deriv pairs", pr0, pr1, "score rel", abs(sc0 - sc1) / abs(sc0), "g rel", np.abs(g0 - g1).max() / np.abs(g0).max(),
          "H rel", np.abs(H0 - H1).max() / np.abs(H0).max())
    G = synth.pose6_to_matrix(gp).astype(np.float32)
    r = O.align(g, prm, of, G)
    t = time.time(); ok, cloud, pose = reg.ScanMatch(of, G); dt = time.time() - t
    lr = reg.last_result
    print("align %.4fs it %d/%d conv %s/%s dp %.3e dT %.3e score %.6f/%.6f" % (dt, lr["iterations"], r["iterations"], lr["converged"],
          r["converged"], np.abs(lr["p"] - r["p"]).max(), np.abs(pose - r["pose"]).max(), lr["score"], r["score"]))
    fit = reg.GetFitnessScore(); ofit = O.fitness_score(m, of, r["pose"])
    print("fitness", fit, ofit, abs(fit - ofit) / ofit)
for C_ in (1, 2, 4, 8, 16):
    reg.SetCluster(C_, 1)
    ts = []
    for rep in range(5):
        t = time.time(); reg.ScanMatch(of, G, want_cloud=False); ts.append(time.time() - t)
    print("cluster", C_, "align ms", np.median(ts) * 1e3, reg.last_result["iterations"])
# batch
B = 64
poses = [synth.pose6_to_matrix(synth.perturb_pose(p, rng)).astype(np.float32) for _ in range(B)]
t = time.time(); P, R = reg.ScanMatchBatch([of] * B, poses); print("batch list", time.time() - t)
t = time.time(); P2, R2 = reg.ScanMatchBatch(of, poses); print("batch shared", time.time() - t, np.array_equal(P, P2))
ok, c1, p1 = reg.ScanMatch(of, poses[5]); print("batch vs single", np.abs(P[5] - p1).max(), R[5]["iterations"], reg.last_result["iterations"])
