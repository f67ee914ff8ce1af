#!/usr/bin/env python
"""Run the BASELINE.json configurations 1, 2, 3 and 5 (config 4 is bench.py) through the reference-facing
API on one B200, beside the CPU oracle, and print one JSON document.

  config 1  scan-to-scan on a synthetic HDL-64 frame pair (res 1.0, step 0.1, eps 0.01, iter 30, VoxelFilter 1.3)
  config 2  front-end odometry: sequential scan-to-local-map NDT, sliding local map of 20 key frames
            (the call pattern of lidar_localization/src/mapping/front_end/front_end.cpp:88-341,348-424)
  config 3  matching node: scan-to-global-map NDT against a 5 M-point map with VoxelFilter 0.6 on the map
            side and a +-100 m crop (src/matching/matching.cpp:148-183,185-265)
  config 5  global relocalisation: 1024 initial-pose hypotheses of one scan scored by NDT, best-fit gather

The oracle runs the same harness on the CPU for a bounded number of frames (parity + baseline timing).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidar_slam_b200 import synth  # noqa: E402
from lidar_slam_b200.callers import FrontEnd, FrontEndDevice  # noqa: E402
from lidar_slam_b200.registration import BoxFilter, DeviceCloud, NDTRegistration, VoxelFilter  # noqa: E402
from oracle import oracle as O  # noqa: E402


def f32(x):
    return float(np.float32(x))


PRM = dict(res=1.0, step_size=0.1, trans_eps=0.01, max_iter=30)


def o_params():
    return O.params(res=1.0, step_size=f32(0.1), trans_eps=f32(0.01), max_iter=30)


def pct(v, q):
    return float(np.percentile(v, q)) if len(v) else None


def pose_err(A, B):
    return float(np.max(np.abs(A[:3, 3] - B[:3, 3]))), float(np.max(np.abs(A[:3, :3] - B[:3, :3])))


# ---------------------------------------------------------------------------------------------- config 1
def config1(scene):
    p0 = scene.path_pose(100.0)
    p1 = p0 + np.array([1.0, 0, 0, 0, 0, np.deg2rad(1.0)])
    f0 = scene.scan(1, p0)
    f1 = scene.scan(2, p1)
    vf = VoxelFilter(1.3, 1.3, 1.3)
    src = vf.Filter(f1)[1]
    reg = NDTRegistration(**PRM)
    t = time.perf_counter(); reg.SetInputTarget(f0); t_tgt = 1e3 * (time.perf_counter() - t)
    grid = O.Grid(f0, 1.0)
    rel = np.linalg.inv(synth.pose6_to_matrix(p0)) @ synth.pose6_to_matrix(p1)
    guesses = {"identity": np.eye(4, dtype=np.float32),
               "truth_perturbed": (rel @ synth.pose6_to_matrix(np.array([0.3, 0, 0, 0, 0, np.deg2rad(1.0)]))).astype(np.float32)}
    out = {"n_target": len(f0), "n_source": len(src), "set_target_ms": t_tgt, "cases": {}}
    for name, g in guesses.items():
        ts = []
        for _ in range(9):
            t = time.perf_counter(); ok, cloud, pose = reg.ScanMatch(src, g, want_cloud=False); ts.append(1e3 * (time.perf_counter() - t))
        t = time.perf_counter(); ref = O.align(grid, o_params(), src, g); t_cpu = 1e3 * (time.perf_counter() - t)
        dt, dr = pose_err(pose, ref["pose"])
        fit, ofit = reg.GetFitnessScore(), O.fitness_score(f0, src, ref["pose"])
        out["cases"][name] = {"gpu_ms_p50": pct(ts, 50), "cpu_oracle_ms": t_cpu, "iterations": reg.last_result["iterations"],
                              "oracle_iterations": ref["iterations"], "pose_dt_m": dt, "pose_dR": dr,
                              "fitness": fit, "fitness_rel_err": abs(fit - ofit) / ofit,
                              "err_vs_truth_m": float(np.max(np.abs(pose[:3, 3] - rel[:3, 3])))}
    return out


# ---------------------------------------------------------------------------------------------- config 2
def config2(scene, frames, oracle_frames):
    s = 60.0 + 1.0 * np.arange(frames)
    truth = np.stack([scene.path_pose(v) for v in s])
    scans = scene.scans(np.arange(frames) + 7000, truth)
    vf, lvf = VoxelFilter(1.3, 1.3, 1.3), VoxelFilter(0.6, 0.6, 0.6)
    reg = NDTRegistration(**PRM)
    fe = FrontEnd(lambda c: vf.Filter(c)[1], lambda c: lvf.Filter(c)[1], reg.SetInputTarget,
                  lambda src, g: reg.ScanMatch(src, g, want_cloud=False)[2])
    T0 = synth.pose6_to_matrix(truth[0])
    traj = [fe.update(scans[k], T0) for k in range(frames)]
    # the same run with device-resident clouds
    fed = FrontEndDevice(VoxelFilter(1.3, 1.3, 1.3), VoxelFilter(0.6, 0.6, 0.6), NDTRegistration(**PRM))
    trajd = [fed.update(scans[k], T0) for k in range(frames)]
    dev_equal = all(np.array_equal(a, b) for a, b in zip(traj, trajd))
    # oracle on the first oracle_frames frames
    state = {}

    def o_set(local):
        state["grid"] = O.Grid(local, 1.0)

    ofe = FrontEnd(lambda c: O.voxel_filter(c, 1.3, 1.3, 1.3)[0], lambda c: O.voxel_filter(c, 0.6, 0.6, 0.6)[0], o_set,
                   lambda src, g: O.align(state["grid"], o_params(), src, g)["pose"])
    otraj = [ofe.update(scans[k], T0) for k in range(min(frames, oracle_frames))]
    dt = max(pose_err(traj[k], otraj[k])[0] for k in range(len(otraj)))
    dR = max(pose_err(traj[k], otraj[k])[1] for k in range(len(otraj)))
    err_truth = [float(np.linalg.norm(traj[k][:3, 3] - synth.pose6_to_matrix(truth[k])[:3, 3])) for k in range(frames)]
    return {"frames": frames, "key_frames": len(fe.t_target), "scan_match_ms": {"p50": pct(fe.t_match, 50), "p99": pct(fe.t_match, 99)},
            "target_rebuild_ms": {"p50": pct(fe.t_target, 50), "max": max(fe.t_target)},
            "local_map_assemble_host_ms_p50": pct(fe.t_assemble, 50),
            "device_resident": {"frame_upload_ms_p50": pct(fed.t_upload, 50),
                                "filter_plus_scan_match_ms": {"p50": pct(fed.t_match, 50), "p99": pct(fed.t_match, 99)},
                                "local_map_assemble_ms_p50": pct(fed.t_assemble, 50),
                                "filter_plus_set_target_ms": {"p50": pct(fed.t_target, 50), "p90": pct(fed.t_target, 90), "max_incl_first_call_module_load": max(fed.t_target)},
                                "trajectory_identical_to_host_path": bool(dev_equal)},
            "oracle": {"frames": len(otraj), "scan_match_ms_p50": pct(ofe.t_match, 50), "target_rebuild_ms_p50": pct(ofe.t_target, 50),
                       "max_traj_dt_m": dt, "max_traj_dR": dR},
            "drift_vs_truth_m": {"final": err_truth[-1], "max": max(err_truth), "every_50_frames": err_truth[::50],
                                 "oracle_final": float(np.linalg.norm(otraj[-1][:3, 3] - synth.pose6_to_matrix(truth[len(otraj) - 1])[:3, 3]))}}


# ---------------------------------------------------------------------------------------------- config 3
def box_crop(cloud, origin, size=100.0):
    """host-side pcl::CropBox with BoxFilter's float edges (box_filter.cpp:63-70)."""
    o = np.asarray(origin, np.float32)
    edge = [np.float32(-size) + o[0], np.float32(size) + o[0], np.float32(-size) + o[1], np.float32(size) + o[1],
            np.float32(-size) + o[2], np.float32(size) + o[2]]
    return O.box_filter(cloud, edge)


def config3(scene, frames, oracle_frames, map_points):
    gmap = scene.make_map(map_points, 2.0)
    t = time.perf_counter(); fmap = VoxelFilter(0.6, 0.6, 0.6).Filter(gmap)[1]; t_mapfilter = 1e3 * (time.perf_counter() - t)
    s = 150.0 + 0.8 * np.arange(frames)
    truth = np.stack([scene.path_pose(v) for v in s])
    scans = scene.scans(np.arange(frames) + 9000, truth)
    vf = VoxelFilter(1.3, 1.3, 1.3)
    reg = NDTRegistration(**PRM)
    origin = truth[0][:3].copy()
    t = time.perf_counter(); local = box_crop(fmap, origin); reg.SetInputTarget(local); t_reset = [1e3 * (time.perf_counter() - t)]
    # device-resident global map: BoxFilter crop + SetInputTarget without host copies
    d_map = DeviceCloud(fmap); d_local = DeviceCloud(); box = BoxFilter([-100.0, 100.0, -100.0, 100.0, -100.0, 100.0])
    reg_d = NDTRegistration(**PRM)
    t_reset_dev, dev_same = [], []

    def reset_dev(org):
        t = time.perf_counter()
        box.SetOrigin(org); box.FilterCloud(d_map, d_local); reg_d.SetInputTargetCloud(d_local)
        t_reset_dev.append(1e3 * (time.perf_counter() - t))
        dev_same.append(bool(reg_d.TargetInfo() == reg.TargetInfo() and len(d_local) == len(local)))

    reset_dev(origin)
    grid = O.Grid(local, 1.0)
    pose = synth.pose6_to_matrix(truth[0] + np.array([0.2, -0.2, 0.05, 0, 0, 0.01])).astype(np.float32)
    last = pose.copy(); predict = pose.copy()
    t_match, t_filter, dts, dRs, errs, o_ms = [], [], [], [], [], []
    for k in range(frames):
        t = time.perf_counter(); src = vf.Filter(scans[k])[1]; t_filter.append(1e3 * (time.perf_counter() - t))
        t = time.perf_counter(); ok, _, pose = reg.ScanMatch(src, predict, want_cloud=False); t_match.append(1e3 * (time.perf_counter() - t))
        if k < oracle_frames:
            t = time.perf_counter(); ref = O.align(grid, o_params(), src, predict); o_ms.append(1e3 * (time.perf_counter() - t))
            a, b = pose_err(pose, ref["pose"]); dts.append(a); dRs.append(b)
        step = np.linalg.inv(last.astype(np.float64)) @ pose.astype(np.float64)
        predict = (pose.astype(np.float64) @ step).astype(np.float32); last = pose.copy()
        errs.append(float(np.linalg.norm(pose[:3, 3] - synth.pose6_to_matrix(truth[k])[:3, 3])))
        if np.any(np.abs(pose[:3, 3] - origin) > 50.0):      # within 50 m of a box edge -> re-crop (matching.cpp:255-262)
            origin = pose[:3, 3].astype(np.float64).copy()
            t = time.perf_counter(); local = box_crop(fmap, origin); reg.SetInputTarget(local); t_reset.append(1e3 * (time.perf_counter() - t))
            reset_dev(origin)
            if k < oracle_frames:
                grid = O.Grid(local, 1.0)
    return {"map_points": len(gmap), "map_filtered": len(fmap), "map_filter_ms": t_mapfilter, "local_map_points": len(local),
            "frames": frames, "frame_filter_ms_p50": pct(t_filter, 50), "scan_match_ms": {"p50": pct(t_match, 50), "p99": pct(t_match, 99)},
            "reset_local_map_ms": t_reset, "reset_local_map_device_resident_ms": t_reset_dev[1:] or t_reset_dev,
            "device_target_identical": all(dev_same), "oracle": {"frames": len(o_ms), "scan_match_ms_p50": pct(o_ms, 50),
                                                      "max_dt_m": max(dts) if dts else None, "max_dR": max(dRs) if dRs else None},
            "err_vs_truth_m": {"p50": pct(errs, 50), "max": max(errs)}}


# ---------------------------------------------------------------------------------------------- config 5
def config5(scene, oracle_hyp):
    target = scene.make_map(1_000_000, 2.0)
    truth = scene.path_pose(420.0)
    scan = scene.scan(555, truth)
    src = VoxelFilter(1.3, 1.3, 1.3).Filter(scan)[1]
    reg = NDTRegistration(**PRM)
    reg.SetInputTarget(target)
    # 32 x 32 lattice of positions, 2 m pitch, centred 3 m off the truth (so no hypothesis is the truth itself)
    gx, gy = np.meshgrid(np.arange(32) - 15.5, np.arange(32) - 15.5, indexing="ij")
    hyp = []
    for dx, dy in zip(gx.ravel(), gy.ravel()):
        p = truth.copy(); p[0] += 2.0 * dx + 0.4; p[1] += 2.0 * dy - 0.3
        hyp.append(synth.pose6_to_matrix(p).astype(np.float32))
    hyp = np.stack(hyp)
    ts = []
    for _ in range(3):
        t = time.perf_counter(); poses, res = reg.ScanMatchBatch(src, hyp); ts.append(1e3 * (time.perf_counter() - t))
    best = int(np.argmax(res["score"]))
    Tt = synth.pose6_to_matrix(truth)
    grid = O.Grid(target, 1.0)
    order = np.argsort(-res["score"])[:oracle_hyp]
    t = time.perf_counter()
    refs = {int(k): O.align(grid, o_params(), src, hyp[k]) for k in order}
    cpu_ms = 1e3 * (time.perf_counter() - t) / max(1, len(order))
    obest = max(refs, key=lambda k: refs[k]["score"])
    dts = [pose_err(poses[k], refs[k]["pose"])[0] for k in refs]
    return {"hypotheses": len(hyp), "n_source": len(src), "wall_ms_all": pct(ts, 50), "hyp_per_s": 1e3 * len(hyp) / pct(ts, 50),
            "best_index": best, "best_err_vs_truth_m": float(np.linalg.norm(poses[best][:3, 3] - Tt[:3, 3])),
            "mean_iterations": float(res["iterations"].mean()),
            "oracle": {"checked_top": len(refs), "cpu_ms_per_hypothesis": cpu_ms, "top1_agrees": bool(obest == best),
                       "max_dt_m": max(dts), "iterations_equal": all(int(res["iterations"][k]) == refs[k]["iterations"] for k in refs)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,5")
    ap.add_argument("--frames2", type=int, default=120)
    ap.add_argument("--frames3", type=int, default=120)
    ap.add_argument("--oracle-frames", type=int, default=25)
    ap.add_argument("--map3", type=int, default=5_000_000)
    args = ap.parse_args()
    O.build(ref=False)
    scene = synth.Scene(leg=500.0)
    out = {}
    for c in args.configs.split(","):
        t = time.time()
        if c == "1": out["config1"] = config1(scene)
        if c == "2": out["config2"] = config2(scene, args.frames2, args.oracle_frames)
        if c == "3": out["config3"] = config3(scene, args.frames3, args.oracle_frames, args.map3)
        if c == "5": out["config5"] = config5(scene, 16)
        out.setdefault("wall_s", {})[c] = time.time() - t
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
