"""-m gpu: NDT target grid, derivative pass, align, fitness through the C ABI vs the oracle.
Tolerances (north_star): voxel assignment + counts bit-exact; pose <= 1e-3 m / 1e-4 rad; fitness <= 1e-4
relative.  Tighter internal gates: per-pass score / gradient / Hessian <= 1e-9 relative at the same
pose, identical iteration counts."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from lidar_slam_b200 import build, capi, synth
from lidar_slam_b200.registration import NDTRegistration, VoxelFilter, to_xyzi8
from tests.conftest import f32

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ndt_small.npz")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pose_close(p_gpu, T_gpu, ref):
    assert np.max(np.abs(T_gpu[:3, 3] - ref["pose"][:3, 3])) <= 1e-3
    assert np.max(np.abs(p_gpu[3:] - ref["p"][3:])) <= 1e-4
    assert np.max(np.abs(T_gpu[:3, :3] - ref["pose"][:3, :3])) <= 1e-4


@pytest.fixture(scope="module")
def setup(oracle, small_map, scans):
    reg = NDTRegistration(1.0, 0.1, 0.01, 30)
    reg.SetInputTarget(small_map)
    grid = oracle.Grid(small_map, 1.0)
    prm = oracle.params(step_size=f32(0.1), trans_eps=f32(0.01))
    srcs = [oracle.voxel_filter(s, 1.3, 1.3, 1.3)[0] for _, s in scans]
    return reg, grid, prm, srcs


def test_target_grid_bit_exact(oracle, setup, small_map):
    reg, grid, prm, _ = setup
    lv = grid.leaves()
    L = reg.TargetLeaves()
    info = reg.TargetInfo()
    lay = grid.layout
    assert info["ok"] and info["min_b"] == list(lay.min_b) and info["div_b"] == list(lay.div_b)
    assert info["n_leaves"] == len(lv) and info["n_tree"] == int((lv["n_raw"] >= 6).sum()) and info["n_points"] == len(small_map)
    assert np.array_equal(L["idx"], lv["idx"]) and np.array_equal(L["n"], lv["n_raw"])
    assert np.array_equal(L["centroid"], lv["centroid"])
    assert np.array_equal(L["mean"], lv["mean"])
    tree = lv["n_raw"] >= 6
    assert np.allclose(L["icov"][tree], lv["icov"][tree], rtol=1e-12, atol=0)
    assert np.mean(np.all(L["icov"][tree] == lv["icov"][tree], axis=1)) > 0.999


def _leaves_equal(A, B):
    return all(np.array_equal(A[k], B[k], equal_nan=True) for k in ("idx", "n", "centroid", "mean", "icov"))


def test_incremental_update_equals_full_build(oracle, small_map, scans):
    """b2ndt_update_target (the in-tree NDT's updateVoxelGrid, VoxelGrid.cpp:545-584,736-809): SetInputTarget(A) followed by
    updates with B1, B2 gives the target SetInputTarget(A ++ B1 ++ B2) gives -- leaf table, matches and fitness bit for bit --
    and that target is the oracle's.  Covers: runs that continue a leaf, runs that open a leaf, crowded runs, a leaf that
    becomes searchable, non-finite points, a cloud outside the index box (rebuild path), host and device clouds."""
    from lidar_slam_b200.registration import DeviceCloud
    n = len(small_map)
    a, b = int(0.6 * n), int(0.85 * n)
    # the points that span the bounding box go first, so the first part already has the index box of the whole cloud
    ext = np.unique(np.concatenate([np.argmin(small_map[:, :3], 0), np.argmax(small_map[:, :3], 0)]))
    small_map = np.concatenate([small_map[ext], np.delete(small_map, ext, 0)])
    A, B1, B2 = small_map[:a], small_map[a:b].copy(), small_map[b:]
    B1[::97, 1] = np.nan                                        # non-finite points are carried along and ignored
    # a dense blob inside one voxel of A's box: a crowded run (> 32 points) that continues / opens a leaf
    rng = np.random.default_rng(5)
    c = A[1000, :3]
    blob = np.concatenate([np.floor(c)[None, :] + rng.uniform(0.05, 0.95, (700, 3)), rng.uniform(0, 1, (700, 1))], 1).astype(np.float32)
    B1 = np.concatenate([B1, blob]).astype(np.float32)
    full = np.concatenate([A, B1, B2]).astype(np.float32)
    ref = NDTRegistration(1.0, 0.1, 0.01, 30)
    ref.SetInputTarget(full)
    want = ref.TargetLeaves()
    lv = oracle.Grid(full, 1.0).leaves()
    assert np.array_equal(want["idx"], lv["idx"]) and np.array_equal(want["n"], lv["n_raw"]) and np.array_equal(want["mean"], lv["mean"])

    src = oracle.voxel_filter(scans[1][1], 1.3, 1.3, 1.3)[0]
    guess = synth.pose6_to_matrix(scans[1][0] + np.array([0.2, -0.15, 0.05, 0.01, -0.005, 0.015])).astype(np.float32)
    _, _, pose_ref = ref.ScanMatch(src, guess)
    res_ref = dict(ref.last_result)
    fit_ref = ref.GetFitnessScore()

    def check(reg, tag):
        info = reg.TargetInfo()
        assert info["n_leaves"] == len(want["idx"]) and info["n_points"] == int(np.isfinite(full[:, :3]).all(1).sum()), tag
        assert info["n_tree"] == ref.TargetInfo()["n_tree"], tag
        assert _leaves_equal(reg.TargetLeaves(), want), tag
        _, _, pose = reg.ScanMatch(src, guess)
        assert np.array_equal(pose, pose_ref), tag
        assert reg.last_result["score"] == res_ref["score"] and reg.last_result["iterations"] == res_ref["iterations"], tag
        assert reg.last_result["pairs"] == res_ref["pairs"], tag
        assert reg.GetFitnessScore() == fit_ref, tag

    # host clouds; A's index box holds B1 and B2, so both updates take the incremental path
    reg = NDTRegistration(1.0, 0.1, 0.01, 30)
    reg.SetInputTarget(A)
    n0 = reg.TargetInfo()["n_leaves"]
    reg.UpdateInputTarget(B1)
    reg.UpdateInputTarget(B2)
    info = reg.TargetInfo()
    assert info["n_leaves"] > n0 and info["updates_incremental"] == 2 and info["updates_rebuilt"] == 0
    check(reg, "host")
    # device clouds
    regd = NDTRegistration(1.0, 0.1, 0.01, 30)
    dA = DeviceCloud(A)
    regd.SetInputTargetCloud(dA)
    regd.UpdateInputTarget(DeviceCloud(B1))
    regd.UpdateInputTarget(DeviceCloud(B2))
    check(regd, "device")
    # rebuild path: the first part lies in a corner of the map, the update reaches outside its index box
    order = np.argsort(full[:, 0], kind="stable")
    lo = np.sort(order[: n // 3]); hi = np.sort(order[n // 3:])
    reg2 = NDTRegistration(1.0, 0.1, 0.01, 30)
    reg2.SetInputTarget(full[lo])
    reg2.UpdateInputTarget(full[hi])
    ref2 = NDTRegistration(1.0, 0.1, 0.01, 30)
    ref2.SetInputTarget(np.concatenate([full[lo], full[hi]]))
    assert _leaves_equal(reg2.TargetLeaves(), ref2.TargetLeaves())
    assert reg2.TargetInfo()["updates_rebuilt"] == 1
    # an update without a target is an error; an empty update is a no-op
    reg3 = NDTRegistration(1.0, 0.1, 0.01, 30)
    with pytest.raises(Exception):
        reg3.UpdateInputTarget(B2)
    reg.UpdateInputTarget(np.zeros((0, 4), np.float32))
    assert _leaves_equal(reg.TargetLeaves(), want)


def test_incremental_update_many_small_updates_grow_the_tables(small_map):
    """A small first target followed by many updates: the leaf tables and the point store are re-allocated several times
    (contents must survive), most updates open new leaves, and the result is still the full build bit for bit.  Also the
    raw-device-pointer entry point (b2ndt_update_target_device) and an update made of non-finite points only."""
    import torch
    ext = np.unique(np.concatenate([np.argmin(small_map[:, :3], 0), np.argmax(small_map[:, :3], 0)]))
    cloud = np.concatenate([small_map[ext], np.delete(small_map, ext, 0)])[:120_000]
    first = 2_000
    reg = NDTRegistration(1.0, 0.1, 0.01, 30)
    reg.SetInputTarget(cloud[:first])
    cuts = np.linspace(first, len(cloud), 15).astype(int)
    for k, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        if k % 3 == 2:
            t = torch.from_numpy(np.ascontiguousarray(cloud[a:b])).cuda()
            capi.check(capi.lib().b2ndt_update_target_device(reg._h, C.c_void_p(t.data_ptr()), b - a))
            reg.Synchronize()
        else:
            reg.UpdateInputTarget(cloud[a:b])
    nan_only = np.full((100, 4), np.nan, np.float32)
    reg.UpdateInputTarget(nan_only)
    info = reg.TargetInfo()
    assert info["updates_incremental"] == 15 and info["updates_rebuilt"] == 0
    ref = NDTRegistration(1.0, 0.1, 0.01, 30)
    ref.SetInputTarget(np.concatenate([cloud, nan_only]))
    assert info["n_points"] == ref.TargetInfo()["n_points"] == len(cloud)
    assert _leaves_equal(reg.TargetLeaves(), ref.TargetLeaves())
    src = cloud[::40].copy()
    guess = synth.pose6_to_matrix(np.array([0.15, -0.1, 0.05, 0.0, 0.0, 0.01])).astype(np.float32)
    _, _, p1 = reg.ScanMatch(src, guess)
    r1 = dict(reg.last_result)
    _, _, p2 = ref.ScanMatch(src, guess)
    assert np.array_equal(p1, p2) and r1["score"] == ref.last_result["score"] and r1["pairs"] == ref.last_result["pairs"]
    assert reg.GetFitnessScore() == ref.GetFitnessScore()
    # a fresh SetInputTarget after updates starts over (counters reset, no stale leaves)
    reg.SetInputTarget(cloud[:first])
    ref.SetInputTarget(cloud[:first])
    assert reg.TargetInfo()["updates_incremental"] == 0 and _leaves_equal(reg.TargetLeaves(), ref.TargetLeaves())


def test_golden_fixture(oracle):
    G = np.load(GOLD)
    reg = NDTRegistration(1.0, 0.1, 0.01, 30)
    reg.SetInputTarget(G["target"])
    L = reg.TargetLeaves()
    assert np.array_equal(L["idx"], G["leaf_idx"]) and np.array_equal(L["n"], G["leaf_n"])
    assert np.array_equal(L["centroid"], G["leaf_centroid"]) and np.array_equal(L["mean"], G["leaf_mean"])
    assert np.allclose(L["icov"], G["leaf_icov"], rtol=1e-12, atol=0)
    for k, q in enumerate(G["deriv_pose"]):
        s, g, H, pairs = reg.Derivatives(G["src"], q)
        assert pairs == G["deriv_pairs"][k]
        assert abs(s - G["deriv_score"][k]) <= 1e-9 * abs(s)
        assert np.allclose(g, G["deriv_grad"][k], rtol=0, atol=1e-9 * np.abs(g).max())
        assert np.allclose(H, G["deriv_hess"][k], rtol=0, atol=1e-9 * np.abs(H).max())
    for compat in (1, 0):
        r2 = NDTRegistration(1.0, 0.1, 0.01, 30, pcl17_compat=bool(compat))
        r2.SetInputTarget(G["target"])
        tag = "c%d_" % compat
        for k, guess in enumerate(G["guesses"]):
            ok, cloud, pose = r2.ScanMatch(G["src"], guess)
            lr = r2.last_result
            assert lr["iterations"] == G[tag + "iterations"][k] and lr["converged"] == G[tag + "converged"][k]
            assert lr["passes"] == G[tag + "passes"][k]
            assert np.max(np.abs(pose[:3, 3] - G[tag + "pose"][k][:3, 3])) <= 1e-3
            assert np.max(np.abs(lr["p"][3:] - G[tag + "p"][k][3:])) <= 1e-4
            assert abs(lr["score"] - G[tag + "score"][k]) <= 1e-6 * max(1.0, abs(G[tag + "score"][k]))
            if compat:
                fit = r2.GetFitnessScore()
                assert abs(fit - G["fitness"][k]) <= 1e-4 * G["fitness"][k]


def test_gpu_matches_reference_intree_outputs():
    """GPU vs outputs of the reference's OWN in-tree NDT code (tests/golden/intree_ndt.npz): per-voxel
    mean bit-identical, inverse covariance to round-off, and the gradient / Hessian of computeDerivatives
    (which the in-tree copy does not weight) to 1e-9."""
    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "intree_ndt.npz"))
    G = np.load(GOLD)
    reg = NDTRegistration(1.0, 0.1, 0.01, 30)
    reg.SetInputTarget(G["target"])
    L = reg.TargetLeaves(); info = reg.TargetInfo()
    tree = L["n"] >= 6
    mb = np.array(info["min_b"]); dv = np.array(info["div_b"])
    idx = L["idx"][tree].astype(np.int64)
    iz = idx // (dv[0] * dv[1]); iy = (idx - iz * dv[0] * dv[1]) // dv[0]; ix = idx - iz * dv[0] * dv[1] - iy * dv[0]
    ijk = np.stack([ix + mb[0], iy + mb[1], iz + mb[2]], 1)
    key = lambda a: a[:, 0].astype(np.int64) * 1_000_000_007 + a[:, 1].astype(np.int64) * 1_000_003 + a[:, 2]
    o = np.argsort(key(ijk)); r = np.argsort(key(ref["vox_ijk"]))
    assert np.array_equal(ijk[o], ref["vox_ijk"][r]) and np.array_equal(L["n"][tree][o], ref["vox_n"][r])
    assert np.array_equal(L["mean"][tree][o], ref["vox_mean"][r])
    a, b = L["icov"][tree][o], ref["vox_icov"][r]
    assert np.max(np.abs(a - b).max(1) / np.abs(b).max(1)) < 1e-12
    for k, q in enumerate(ref["deriv_pose"]):
        s, g, H, pairs = reg.Derivatives(G["src"], q)
        assert np.max(np.abs(g - ref["deriv_grad"][k])) <= 1e-9 * np.abs(ref["deriv_grad"][k]).max()
        assert np.max(np.abs(H - ref["deriv_hess"][k])) <= 1e-9 * np.abs(ref["deriv_hess"][k]).max()
    for k, guess in enumerate(ref["guesses"]):
        ok, cloud, pose = reg.ScanMatch(G["src"], guess)
        assert reg.last_result["iterations"] == ref["align_iterations"][k]
        assert np.max(np.abs(pose[:3, 3] - ref["align_pose"][k][:3, 3])) <= 1e-3
        assert np.max(np.abs(pose[:3, :3] - ref["align_pose"][k][:3, :3])) <= 1e-4


def test_derivative_pass_matches_oracle(oracle, setup, scans):
    reg, grid, prm, srcs = setup
    rng = np.random.default_rng(3)
    for (truth, _), src in zip(scans, srcs):
        for C_ in (1, 4, 16):
            reg.SetCluster(C_, 1)
            q = synth.perturb_pose(truth, rng)
            s0, g0, H0, p0 = oracle.derivatives(grid, prm, src, q)
            s1, g1, H1, p1 = reg.Derivatives(src, q)
            assert p0 == p1, "neighbour sets differ"
            assert abs(s0 - s1) <= 1e-9 * abs(s0)
            assert np.max(np.abs(g0 - g1)) <= 1e-9 * np.abs(g0).max()
            assert np.max(np.abs(H0 - H1)) <= 1e-9 * np.abs(H0).max()
    reg.SetCluster(8, 1)


def test_align_matches_oracle(oracle, setup, scans, small_map):
    reg, grid, prm, srcs = setup
    rng = np.random.default_rng(4)
    for (truth, _), src in zip(scans, srcs):
        for trial in range(3):
            guess = synth.pose6_to_matrix(synth.perturb_pose(truth, rng)).astype(np.float32)
            ref = oracle.align(grid, prm, src, guess, want_cloud=True)
            ok, cloud, pose = reg.ScanMatch(src, guess)
            lr = reg.last_result
            assert ok and lr["iterations"] == ref["iterations"] and lr["converged"] == ref["converged"]
            assert lr["passes"] == ref["passes"] and lr["pairs"] == ref["pairs"]
            _pose_close(lr["p"], pose, ref)
            assert abs(lr["score"] - ref["score"]) <= 1e-6 * abs(ref["score"])
            assert abs(lr["trans_probability"] - ref["trans_probability"]) <= 1e-6 * abs(ref["trans_probability"])
            assert np.max(np.abs(cloud[:, :3] - ref["cloud"])) <= 2e-3
            fit = reg.GetFitnessScore()
            ofit = oracle.fitness_score(small_map, src, ref["pose"])
            assert abs(fit - ofit) <= 1e-4 * ofit
    # identity guess (the `guess != Identity` branch) and PointXYZI layout
    src = srcs[0]
    ident = np.eye(4, dtype=np.float32)
    ref = oracle.align(grid, prm, src, ident)
    ok, cloud, pose = reg.ScanMatch(to_xyzi8(src), ident)
    assert reg.last_result["iterations"] == ref["iterations"]
    _pose_close(reg.last_result["p"], pose, ref)


def test_result_cloud_filled_on_device_equals_host_transform(setup, scans):
    """b2ndt_align_ex: the result cloud written from the device copy of the source == the source moved by the final pose
    with pcl::transformPointCloud's float arithmetic (what the host loop of the drop-in class computes)."""
    reg, grid, prm, srcs = setup
    guess = synth.pose6_to_matrix(scans[0][0] + np.array([0.1, 0.1, 0.0, 0.0, 0.0, 0.01])).astype(np.float32)
    ok, c_host, pose_h = reg.ScanMatch(srcs[0], guess)
    ok, c_dev, pose_d = reg.ScanMatch(srcs[0], guess, want_cloud="device")
    assert np.array_equal(pose_h, pose_d)
    assert c_dev.shape == (len(srcs[0]), 4) and np.array_equal(c_dev[:, 3], srcs[0][:, 3])
    assert np.array_equal(c_dev[:, :3], c_host[:, :3])
    # PointXYZI records (32 bytes, intensity at +16, data[3] = 1), written over the source buffer itself
    src8 = to_xyzi8(srcs[0])
    buf = src8.copy()
    out = np.zeros(16, np.float32)
    res = capi.Result()
    g = capi.pose_to_colmajor(guess)
    capi.check(capi.lib().b2ndt_align_ex(reg._h, buf.ctypes.data_as(C.c_void_p), len(buf), 32, 16, capi._fp(g), capi._fp(out), C.byref(res),
                                         buf.ctypes.data_as(C.c_void_p), 32, 16))
    assert np.array_equal(buf[:, :3], c_host[:, :3]) and np.all(buf[:, 3] == 1.0) and np.array_equal(buf[:, 4], srcs[0][:, 3])
    assert np.all(buf[:, 5:] == 0.0)


def test_cluster_widths_agree(oracle, setup, scans, small_map):
    reg, grid, prm, srcs = setup
    truth, _ = scans[1]
    guess = synth.pose6_to_matrix(truth + np.array([0.3, -0.2, 0.1, 0.01, 0.01, -0.02])).astype(np.float32)
    res = []
    for C_ in (1, 2, 4, 8, 16):
        reg.SetCluster(C_, 1)
        ok, _, pose = reg.ScanMatch(srcs[1], guess, want_cloud=False)
        res.append((pose, reg.last_result["iterations"]))
    reg.SetCluster(8, 1)
    for pose, it in res[1:]:
        assert it == res[0][1] and np.max(np.abs(pose - res[0][0])) <= 1e-6
    # run-to-run reproducibility (fixed-order reductions)
    ok, _, p1 = reg.ScanMatch(srcs[1], guess, want_cloud=False)
    ok, _, p2 = reg.ScanMatch(srcs[1], guess, want_cloud=False)
    assert np.array_equal(p1, p2)
    # the two CTA shapes of ndt_match_kernel: a launch with at most one CTA per SM runs 8 compute + 8 search warps per CTA
    # (<8>), B2NDT_WIDE_SINGLE=0 (read when a handle is created) keeps the 4 + 8 shape (<4>) that larger launches use
    os.environ["B2NDT_WIDE_SINGLE"] = "0"
    try:
        narrow = NDTRegistration(1.0, 0.1, 0.01, 30)
    finally:
        del os.environ["B2NDT_WIDE_SINGLE"]
    narrow.SetInputTarget(small_map)
    for C_ in (1, 8, 11):
        narrow.SetCluster(C_, 1)
        ok, _, pose = narrow.ScanMatch(srcs[1], guess, want_cloud=False)
        assert ok and narrow.last_result["iterations"] == res[0][1] and np.max(np.abs(pose - res[0][0])) <= 1e-6
        assert narrow.last_result["pairs"] == reg.last_result["pairs"]


def test_more_thuente_enabled_matches_oracle(oracle, small_map, scans):
    reg = NDTRegistration(1.0, 0.1, 0.01, 30, pcl17_compat=False)
    reg.SetInputTarget(small_map)
    grid = oracle.Grid(small_map, 1.0)
    prm = oracle.params(step_size=f32(0.1), trans_eps=f32(0.01), pcl17_compat=0)
    rng = np.random.default_rng(5)
    trials = 0
    for truth, scan in scans[:3]:
        src = oracle.voxel_filter(scan, 1.3, 1.3, 1.3)[0]
        for _ in range(2):
            guess = synth.pose6_to_matrix(synth.perturb_pose(truth, rng)).astype(np.float32)
            ref = oracle.align(grid, prm, src, guess)
            ok, _, pose = reg.ScanMatch(src, guess, want_cloud=False)
            lr = reg.last_result
            assert lr["iterations"] == ref["iterations"] and lr["mt_trials"] == ref["mt_trials"] and lr["passes"] == ref["passes"]
            _pose_close(lr["p"], pose, ref)
            trials += ref["mt_trials"]
    assert trials > 0, "the line search never ran"


def test_batch_equals_single_matches(oracle, setup, scans):
    reg, grid, prm, srcs = setup
    rng = np.random.default_rng(6)
    B = 48
    sources, guesses = [], []
    for b in range(B):
        k = b % len(srcs)
        sources.append(srcs[k][: len(srcs[k]) - (b % 7)])     # ragged lengths
        guesses.append(synth.pose6_to_matrix(synth.perturb_pose(scans[k][0], rng)).astype(np.float32))
    poses, res = reg.ScanMatchBatch(sources, guesses)
    for b in (0, 5, 17, 47):
        ok, _, p1 = reg.ScanMatch(sources[b], guesses[b], want_cloud=False)
        assert np.max(np.abs(p1 - poses[b])) <= 1e-6 and reg.last_result["iterations"] == res[b]["iterations"]
        ref = oracle.align(grid, prm, sources[b], guesses[b])
        assert res[b]["iterations"] == ref["iterations"]
        _pose_close(res[b]["p"], poses[b], ref)
    # cluster of 2 CTAs per match in batch mode gives the same answers
    reg.SetCluster(8, 2)
    poses2, res2 = reg.ScanMatchBatch(sources, guesses)
    reg.SetCluster(8, 1)
    assert np.max(np.abs(poses2 - poses)) <= 1e-6 and np.array_equal(res2["iterations"], res["iterations"])
    # relocalisation hypotheses: one source x B guesses (config 5)
    hyp = [synth.pose6_to_matrix(synth.perturb_pose(scans[0][0], rng, 1.0, 3.0)).astype(np.float32) for _ in range(32)]
    ph, rh = reg.ScanMatchBatch(srcs[0], hyp)
    pl, rl = reg.ScanMatchBatch([srcs[0]] * 32, hyp)
    assert np.array_equal(ph, pl) and np.array_equal(rh["iterations"], rl["iterations"])
    # empty batch member
    sources[3] = np.zeros((0, 4), np.float32)
    poses3, res3 = reg.ScanMatchBatch(sources, guesses)
    assert res3[3]["iterations"] == 0 and np.max(np.abs(poses3[4] - poses[4])) <= 1e-6


def test_degenerate_inputs(oracle, small_map):
    reg = NDTRegistration(1.0, 0.1, 0.01, 30)
    from lidar_slam_b200 import capi
    with pytest.raises(capi.B2Error) as e:
        reg.ScanMatch(np.zeros((4, 4), np.float32), np.eye(4, dtype=np.float32))
    assert e.value.code == capi.B2_ERR_STATE
    reg.SetInputTarget(small_map)
    with pytest.raises(capi.B2Error):
        reg.GetFitnessScore()                       # no ScanMatch yet
    far = np.array([[1e4, 1e4, 50, 0], [1e4 + 1, 1e4, 50, 0]], np.float32)
    ok, cloud, pose = reg.ScanMatch(far, np.eye(4, dtype=np.float32))
    assert reg.last_result["converged"] and reg.last_result["iterations"] == 0 and np.array_equal(pose, np.eye(4, dtype=np.float32))
    # far-away queries exercise the fitness brute-force fallback
    fit = reg.GetFitnessScore()
    ofit = oracle.fitness_score(small_map, far, np.eye(4, dtype=np.float32))
    assert abs(fit - ofit) <= 1e-4 * ofit
    # a derivative pass or a batch reuses the handle's source buffer: the "last ScanMatch" is gone
    reg.Derivatives(far, np.zeros(6))
    with pytest.raises(capi.B2Error) as e:
        reg.GetFitnessScore()
    assert e.value.code == capi.B2_ERR_STATE
    # empty target: PCL leaves no cells
    reg.SetInputTarget(np.zeros((0, 4), np.float32))
    ok, cloud, pose = reg.ScanMatch(far, np.eye(4, dtype=np.float32))
    assert reg.last_result["iterations"] == 0
    # target with NaNs and fewer than 6 points per voxel anywhere
    t = np.array([[0, 0, 0, 1], [np.nan, 1, 1, 1], [5, 5, 5, 1]], np.float32)
    reg.SetInputTarget(t)
    info = reg.TargetInfo()
    assert info["n_points"] == 2 and info["n_leaves"] == 2 and info["n_tree"] == 0


def test_fitness_matches_oracle(oracle, setup, small_map, scans):
    reg, grid, prm, srcs = setup
    for (truth, scan), src in zip(scans[:2], srcs[:2]):
        T = synth.pose6_to_matrix(truth + np.array([0.05, 0.02, 0, 0, 0, 0.002])).astype(np.float32)
        got = reg.GetFitnessScoreFor(src, T)
        want = oracle.fitness_score(small_map, src, T)
        assert abs(got - want) <= 1e-9 * want
        got = reg.GetFitnessScoreFor(scan[::5], T, max_range=0.5)
        want = oracle.fitness_score(small_map, scan[::5], T, max_range=0.5)
        assert abs(got - want) <= 1e-9 * want


def test_cpp_dropin_classes(oracle, small_map, scans, tmp_path):
    """C++ NDTRegistration / VoxelFilter (reference interface) driven like front_end.cpp / matching.cpp."""
    build.build_host()
    src = os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp")
    exe = str(tmp_path / "test_dropin")
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-I", os.path.join(ROOT, "include"), "-o", exe, src,
                           "-L", build.LIBDIR, "-lb2host", "-lb2ndt", "-Wl,-rpath," + build.LIBDIR])
    truth, scan = scans[0]
    guess = synth.pose6_to_matrix(truth + np.array([0.2, 0.1, -0.1, 0.01, -0.01, 0.015])).astype(np.float32)
    small_map.tofile(str(tmp_path / "t.bin")); scan.tofile(str(tmp_path / "s.bin"))
    np.ascontiguousarray(guess.flatten(order="F")).tofile(str(tmp_path / "g.bin"))
    out = subprocess.run([exe, str(tmp_path / "t.bin"), str(tmp_path / "s.bin"), str(tmp_path / "g.bin")],
                         stdout=subprocess.PIPE, text=True, check=True).stdout
    kv = {l.split()[0]: l.split()[1:] for l in out.splitlines() if l and l.split()[0] in ("M", "POSE", "FIT", "ITER", "R0", "BOX", "DEV", "UPD")}
    assert kv["DEV"][0] == "1" and kv["DEV"][1] == kv["ITER"][0]          # device-resident C++ path == host-buffer path
    assert kv["UPD"][0] == "1" and kv["UPD"][1] == kv["UPD"][2]           # UpdateInputTarget == SetInputTarget of the whole; device-filled result cloud
    assert kv["BOX"][0] == kv["BOX"][1] and int(kv["BOX"][0]) > 0 and kv["BOX"][2] == kv["BOX"][3]      # C++ BoxFilter == CropBox
    filt = oracle.voxel_filter(scan, 1.3, 1.3, 1.3)[0]
    grid = oracle.Grid(small_map, 1.0)
    ref = oracle.align(grid, oracle.params(step_size=f32(0.1), trans_eps=f32(0.01)), filt, guess, want_cloud=True)
    assert int(kv["M"][0]) == len(filt) and int(kv["ITER"][0]) == ref["iterations"]
    pose = np.array([float(v) for v in kv["POSE"]], np.float32).reshape(4, 4, order="F")
    assert np.max(np.abs(pose[:3, 3] - ref["pose"][:3, 3])) <= 1e-3 and np.max(np.abs(pose[:3, :3] - ref["pose"][:3, :3])) <= 1e-4
    ofit = oracle.fitness_score(small_map, filt, ref["pose"])
    assert abs(float(kv["FIT"][0]) - ofit) <= 1e-4 * ofit
    r0 = np.array([float(v) for v in kv["R0"]])
    assert np.max(np.abs(r0[:3] - ref["cloud"][0])) <= 2e-3 and abs(r0[3] - filt[0, 3]) < 1e-6


@pytest.mark.gpu
def test_streamed_host_batch_equals_resident_batch(setup, scans, small_map):
    """Large host batches run as ONE persistent launch with the sources streaming in behind it (16 chunks, a
    counter of resident matches): same poses / iterations as the same batch with everything copied first, for
    packed pinned-or-not sources and for the 32-byte PointXYZI layout (host repack per chunk)."""
    import os
    from lidar_slam_b200.registration import NDTRegistration, to_xyzi8
    reg, grid, prm, srcs = setup
    rng = np.random.default_rng(16)
    B = 300
    sources, guesses = [], []
    for b in range(B):
        k = b % len(srcs)
        sources.append(srcs[k][: len(srcs[k]) - (b % 11)])
        guesses.append(synth.pose6_to_matrix(synth.perturb_pose(scans[k][0], rng)).astype(np.float32))
    sources[7] = np.zeros((0, 4), np.float32)                      # an empty member inside a chunk
    poses, res = reg.ScanMatchBatch(sources, guesses)              # streamed (B >= 256)
    poses8, res8 = reg.ScanMatchBatch([to_xyzi8(s) for s in sources], guesses)
    os.environ["B2NDT_STREAM"] = "0"
    try:
        reg0 = NDTRegistration(1.0, 0.1, 0.01, 30)
        reg0.SetInputTarget(small_map)
        poses0, res0 = reg0.ScanMatchBatch(sources, guesses)
    finally:
        del os.environ["B2NDT_STREAM"]
    assert np.array_equal(poses, poses0) and np.array_equal(res["iterations"], res0["iterations"])
    assert np.array_equal(poses8, poses0) and np.array_equal(res8["pairs"], res0["pairs"])
    assert res[7]["iterations"] == 0


@pytest.mark.gpu
def test_batch_kernel_edge_cases(setup, scans, small_map):
    """Persistent two-slot batch kernel: tiny batches (a slot with no work), far-off guesses that run to the iteration
    cap, all-empty sources and a target without searchable voxels give the same answers as single matches."""
    reg, grid, prm, srcs = setup
    rng = np.random.default_rng(23)
    guesses = [synth.pose6_to_matrix(synth.perturb_pose(scans[k % len(srcs)][0], rng)).astype(np.float32) for k in range(5)]
    far = synth.pose6_to_matrix(scans[0][0] + np.array([6.0, -5.0, 1.0, 0.05, -0.04, 0.6])).astype(np.float32)
    for B in (2, 3, 5):
        sources = [srcs[k % len(srcs)] for k in range(B)]
        g = list(guesses[:B])
        g[B - 1] = far
        poses, res = reg.ScanMatchBatch(sources, g)
        for b in range(B):
            ok, _, p1 = reg.ScanMatch(sources[b], g[b], want_cloud=False)
            assert np.array_equal(p1, poses[b]), (B, b)
            assert reg.last_result["iterations"] == res[b]["iterations"] and reg.last_result["pairs"] == res[b]["pairs"]
    # every source empty
    poses, res = reg.ScanMatchBatch([np.zeros((0, 4), np.float32)] * 4, guesses[:4])
    assert all(r["iterations"] == 0 for r in res)
    # a target whose voxels all hold fewer than min_points_per_voxel points: no pairs anywhere
    sparse = small_map[::40].copy()
    reg2 = NDTRegistration(0.3, 0.1, 0.01, 30)
    reg2.SetInputTarget(sparse)
    assert reg2.TargetInfo()["n_tree"] == 0
    poses, res = reg2.ScanMatchBatch([srcs[0], srcs[1], srcs[0]], guesses[:3])
    for b in range(3):
        ok, _, p1 = reg2.ScanMatch([srcs[0], srcs[1], srcs[0]][b], guesses[b], want_cloud=False)
        assert np.array_equal(p1, poses[b]) and res[b]["pairs"] == 0


@pytest.mark.gpu
def test_one_process_several_handles_in_threads(setup, scans, small_map):
    """one host thread + handle + stream per device from ONE process.  On a single-GPU box the handles
    share device 0; results equal the single-handle batch, so handles are independent and the library is safe to
    drive from several threads (one handle per thread)."""
    import torch
    from lidar_slam_b200 import batch
    reg, grid, prm, srcs = setup
    ndev = max(1, torch.cuda.device_count())
    regs = []
    for r in range(3):
        h = NDTRegistration(1.0, 0.1, 0.01, 30, device=r % ndev)
        h.SetInputTarget(small_map)
        regs.append(h)
    rng = np.random.default_rng(31)
    B = 41
    sources = [srcs[b % len(srcs)][: len(srcs[b % len(srcs)]) - (b % 5)] for b in range(B)]
    guesses = [synth.pose6_to_matrix(synth.perturb_pose(scans[b % len(srcs)][0], rng)).astype(np.float32) for b in range(B)]
    poses, res = batch.scan_match_batch_multi_device(regs, sources, guesses)
    poses1, res1 = reg.ScanMatchBatch(sources, guesses)
    assert np.array_equal(poses, poses1) and np.array_equal(res["iterations"], res1["iterations"])
    assert np.array_equal(res["pairs"], res1["pairs"])
    # one shared source x B hypotheses (config 5), split over the handles
    ph, rh = batch.scan_match_batch_multi_device(regs, srcs[0], guesses)
    p1, r1 = reg.ScanMatchBatch(srcs[0], guesses)
    assert np.array_equal(ph, p1) and np.array_equal(rh["iterations"], r1["iterations"])
