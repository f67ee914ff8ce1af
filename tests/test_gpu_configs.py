"""BASELINE.json configs 2, 3 and 5 replayed through the reference-facing API on the GPU, every frame / hypothesis
checked against the CPU oracle driving the SAME caller loop (lidar_slam_b200/callers.py):

  config 2  front_end.cpp:88-341,348-424   sequential scan-to-local-map odometry, sliding local map of 20 key frames
  config 3  matching.cpp:148-183,185-265   5 M-point map -> VoxelFilter 0.6 -> +-100 m crop -> ScanMatch, with re-crops
  config 5  matching.cpp:267-308 generalised  1024 initial-pose hypotheses of one scan, best fit

Tolerances are north_star's: pose 1e-3 m / 1e-4 rad, iterations equal; voxel ids / counts / centroids bit-exact."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from lidar_slam_b200 import synth
from lidar_slam_b200.callers import FrontEnd, FrontEndDevice, MatchingLoop, box_edges, hypothesis_lattice
from oracle import oracle as O

pytestmark = pytest.mark.gpu

PRM = dict(res=1.0, step_size=0.1, trans_eps=0.01, max_iter=30)


def o_params():
    return O.params(res=1.0, step_size=float(np.float32(0.1)), trans_eps=float(np.float32(0.01)), max_iter=30)


def rot_err(A, B):
    """rotation difference in radians between two 4x4 float poses: |R_a - R_b|_F / sqrt(2) (= the angle for small
    differences; arccos of the trace cannot resolve 1e-4 rad on float32 matrices)"""
    return float(np.linalg.norm(A[:3, :3].astype(np.float64) - B[:3, :3].astype(np.float64)) / np.sqrt(2.0))


def test_config2_front_end_trajectory_equals_oracle_front_end():
    """>= 60 frames, >= 25 key-frame target rebuilds (incl. the switch to the filtered local map at 10 key frames and
    the sliding window once 20 exist): host-buffer path, device-resident path and oracle front end, frame by frame."""
    from lidar_slam_b200.registration import NDTRegistration, VoxelFilter
    O.build(ref=False)
    scene = synth.Scene(leg=500.0)
    frames = 84
    s = 60.0 + 1.0 * np.arange(frames)
    truth = np.stack([scene.path_pose(v) for v in s])
    scans = scene.scans(np.arange(frames) + 7000, truth, nthreads=max(1, (os.cpu_count() or 2) // 2))
    T0 = synth.pose6_to_matrix(truth[0])

    vf, lvf = VoxelFilter(1.3, 1.3, 1.3), VoxelFilter(0.6, 0.6, 0.6)
    reg = NDTRegistration(**PRM)
    fe = FrontEnd(lambda c: vf.Filter(c)[1], lambda c: lvf.Filter(c)[1], reg.SetInputTarget,
                  lambda src, g: reg.ScanMatch(src, g, want_cloud=False)[2])
    traj = [fe.update(scans[k], T0) for k in range(frames)]

    fed = FrontEndDevice(VoxelFilter(1.3, 1.3, 1.3), VoxelFilter(0.6, 0.6, 0.6), NDTRegistration(**PRM))
    trajd = [fed.update(scans[k], T0) for k in range(frames)]

    state = {}
    ofe = FrontEnd(lambda c: O.voxel_filter(c, 1.3, 1.3, 1.3)[0], lambda c: O.voxel_filter(c, 0.6, 0.6, 0.6)[0],
                   lambda local: state.__setitem__("grid", O.Grid(local, 1.0)),
                   lambda src, g: O.align(state["grid"], o_params(), src, g)["pose"], transform=O.transform_cloud)
    otraj = [ofe.update(scans[k], T0) for k in range(frames)]

    assert len(fe.t_target) >= 25 and len(fe.t_target) == len(ofe.t_target) == len(fed.t_target), len(fe.t_target)
    assert len(fe.keyframes) == 20                      # the window slid
    for k in range(frames):
        assert np.max(np.abs(traj[k][:3, 3] - otraj[k][:3, 3])) <= 1e-3 and rot_err(traj[k], otraj[k]) <= 1e-4, k
        # device-resident clouds: the same kernels on the same data
        assert np.array_equal(traj[k], trajd[k]), k
    # odometry sanity against the ground truth (the street canyon is weakly constrained along the street)
    lateral = [abs(float(traj[k][1, 3] - synth.pose6_to_matrix(truth[k])[1, 3])) for k in range(frames)]
    assert np.isfinite(lateral).all() and max(lateral) < 2.0, max(lateral)


def test_config3_map_matching_with_recrops_every_frame_vs_oracle():
    """5 M-point map -> VoxelFilter 0.6 (bit-exact vs the oracle filter) -> BoxFilter +-100 m -> SetInputTarget; 125
    frames with >= 2 re-crops; every frame's pose vs the oracle running the same loop on the same filtered map; the
    device-resident crop + target build gives the same target as the host path at every re-crop."""
    from lidar_slam_b200.registration import BoxFilter, DeviceCloud, NDTRegistration, VoxelFilter
    O.build(ref=False)
    scene = synth.Scene(leg=500.0)
    gmap = scene.make_map(5_000_000, 2.0)
    ok, fmap, idx, cnt = VoxelFilter(0.6, 0.6, 0.6).Filter(gmap, with_info=True)
    o_fmap, o_idx, o_cnt, _ = O.voxel_filter(gmap, 0.6, 0.6, 0.6)
    assert np.array_equal(idx, o_idx) and np.array_equal(cnt, o_cnt) and np.array_equal(fmap, o_fmap)

    frames = 125
    s = 150.0 + 1.0 * np.arange(frames)                 # 125 m of driving: the +-100 m box is re-centred every ~50 m
    truth = np.stack([scene.path_pose(v) for v in s])
    scans = scene.scans(np.arange(frames) + 9000, truth, nthreads=max(1, (os.cpu_count() or 2) // 2))
    vf = VoxelFilter(1.3, 1.3, 1.3)
    reg = NDTRegistration(**PRM)
    d_map, d_local = DeviceCloud(fmap), DeviceCloud()
    box = BoxFilter([-100.0, 100.0, -100.0, 100.0, -100.0, 100.0])
    reg_d = NDTRegistration(**PRM)
    dev_same = []

    def crop_host(origin):
        local = O.box_filter(fmap, box_edges(origin))
        # the device-resident re-crop must produce the same cloud and the same target
        box.SetOrigin(origin); box.FilterCloud(d_map, d_local); reg_d.SetInputTargetCloud(d_local)
        dev_same.append(len(d_local) == len(local) and np.array_equal(d_local.Download(), local))
        return local

    def set_target(local):
        reg.SetInputTarget(local)
        dev_same.append(reg_d.TargetInfo() == reg.TargetInfo())

    loop = MatchingLoop(lambda c: vf.Filter(c)[1], crop_host, set_target, lambda src, g: reg.ScanMatch(src, g, want_cloud=False)[2])
    state = {}
    oloop = MatchingLoop(lambda c: O.voxel_filter(c, 1.3, 1.3, 1.3)[0], lambda origin: O.box_filter(fmap, box_edges(origin)),
                         lambda local: state.__setitem__("grid", O.Grid(local, 1.0)),
                         lambda src, g: O.align(state["grid"], o_params(), src, g)["pose"])
    init = synth.pose6_to_matrix(truth[0] + np.array([0.2, -0.2, 0.05, 0, 0, 0.01])).astype(np.float32)
    loop.set_init_pose(init); oloop.set_init_pose(init)
    for k in range(frames):
        pose, opose = loop.update(scans[k]), oloop.update(scans[k])
        assert np.max(np.abs(pose[:3, 3] - opose[:3, 3])) <= 1e-3 and rot_err(pose, opose) <= 1e-4, k
        assert loop.recrops == oloop.recrops, k
    assert loop.recrops >= 2, loop.recrops
    assert all(dev_same)
    err = np.linalg.norm(pose[:3, 3] - synth.pose6_to_matrix(truth[-1])[:3, 3])
    assert err < 1.5, err


def test_config5_1024_hypotheses_topk_equals_oracle():
    """One scan x 1024 hypotheses (32 x 32 lattice, 2 m pitch) in ONE batched launch with a shared source: every
    hypothesis' iteration count and final pose vs the oracle, and the ranking by score (top 16) identical."""
    from lidar_slam_b200.registration import NDTRegistration, VoxelFilter
    O.build(ref=False)
    scene = synth.Scene(leg=500.0)
    target = scene.make_map(1_000_000, 2.0)
    truth = scene.path_pose(420.0)
    src = VoxelFilter(1.3, 1.3, 1.3).Filter(scene.scan(555, truth))[1]
    hyp = hypothesis_lattice(truth, synth.pose6_to_matrix)
    assert len(hyp) == 1024
    reg = NDTRegistration(**PRM)
    reg.SetInputTarget(target)
    poses, res = reg.ScanMatchBatch(src, hyp)
    grid = O.Grid(target, 1.0)
    prm = o_params()
    with ThreadPoolExecutor(max_workers=max(1, os.cpu_count() or 1)) as ex:
        refs = list(ex.map(lambda k: O.align(grid, prm, src, hyp[k]), range(len(hyp))))
    o_score = np.array([r["score"] for r in refs])
    assert np.array_equal(res["iterations"], np.array([r["iterations"] for r in refs]))
    dt = max(float(np.max(np.abs(poses[k][:3, 3] - refs[k]["pose"][:3, 3]))) for k in range(len(hyp)))
    dr = max(float(np.max(np.abs(res["p"][k][3:] - refs[k]["p"][3:]))) for k in range(len(hyp)))
    assert dt <= 1e-3 and dr <= 1e-4, (dt, dr)
    assert np.allclose(res["score"], o_score, rtol=1e-6, atol=1e-9)      # final score after ~20 iterations: summation order
    top, otop = np.argsort(-res["score"], kind="stable")[:16], np.argsort(-o_score, kind="stable")[:16]
    assert np.array_equal(top, otop)
    best = int(top[0])
    assert np.linalg.norm(poses[best][:3, 3] - synth.pose6_to_matrix(truth)[:3, 3]) < 0.05
    # sharded over "ranks" (what bench.py --workload config5 does across GPUs): same answers, same winner
    from lidar_slam_b200 import batch
    parts = [reg.ScanMatchBatch(src, hyp[lo:hi]) for lo, hi in (batch.shard_range(len(hyp), r, 8) for r in range(8))]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), poses)
    rows = np.concatenate([p[1]["score"] for p in parts])
    assert int(np.argmax(rows)) == best
