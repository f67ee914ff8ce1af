"""Device-resident clouds: local-map assembly, BoxFilter crop, VoxelFilter and
SetInputTarget / ScanMatch without leaving HBM, checked against the oracle and against the host-buffer path."""
import numpy as np
import pytest

from lidar_slam_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    return synth.Scene(leg=60.0)


def test_upload_download_roundtrip(scene):
    from lidar_slam_b200.registration import DeviceCloud, to_xyzi8
    scan = scene.scan(3, scene.path_pose(10.0))
    dc = DeviceCloud(scan)
    assert len(dc) == len(scan)
    assert np.array_equal(dc.Download(), scan)
    assert np.array_equal(dc.Download(layout=8), to_xyzi8(scan))
    dc8 = DeviceCloud(to_xyzi8(scan))
    assert np.array_equal(dc8.Download(), scan)
    empty = DeviceCloud(np.zeros((0, 4), np.float32))
    assert len(empty) == 0 and empty.Download().shape == (0, 4)


def test_box_filter_matches_cropbox_oracle(scene):
    from lidar_slam_b200.registration import BoxFilter, DeviceCloud
    m = scene.make_map(200000, 2.0)
    m[17, 0] = np.nan
    m[99, 2] = np.inf
    bf = BoxFilter({"box_filter_size": [-30.0, 30.0, -25.0, 25.0, -1.0, 6.0]})
    bf.SetOrigin([10.0, -5.0, 0.5])
    edge = bf.GetEdge()
    assert edge == [-20.0, 40.0, -30.0, 20.0, -0.5, 6.5]
    ok, out = bf.Filter(m)
    ref = O.box_filter(m, edge)
    assert ok and out.shape == ref.shape and np.array_equal(out, ref)      # same points, same order
    # boundary is inclusive (pcl::CropBox rejects only < min or > max)
    pts = np.array([[edge[0], edge[2], edge[4], 1.0], [edge[1], edge[3], edge[5], 2.0],
                    [np.nextafter(np.float32(edge[0]), np.float32(-1e9)), 0, 0, 3.0]], np.float32)
    ok, out = bf.Filter(pts)
    assert np.array_equal(out, pts[:2])
    # empty input / nothing kept
    assert bf.Filter(np.zeros((0, 4), np.float32))[1].shape == (0, 4)
    far = m.copy(); far[:, 0] += 1e4
    assert len(bf.FilterCloud(DeviceCloud(far))) == 0


def test_local_map_assembly_filter_and_target(scene):
    """front_end.cpp:375-410 on the device: sum of transformed key frames -> VoxelFilter -> SetInputTarget;
    bit-identical to the same steps through host buffers, and the match equals the oracle's."""
    from lidar_slam_b200.registration import DeviceCloud, NDTRegistration, VoxelFilter
    rng = np.random.default_rng(5)
    frames, poses = [], []
    for k in range(6):
        pose6 = scene.path_pose(8.0 + 2.0 * k)
        frames.append(scene.scan(100 + k, pose6)[::3].copy())
        poses.append(synth.pose6_to_matrix(pose6).astype(np.float32))
    inv0 = np.linalg.inv(poses[0].astype(np.float64)).astype(np.float32)
    rel = [(inv0 @ P).astype(np.float32) for P in poses]
    # device
    local = DeviceCloud()
    for f, T in zip(frames, rel):
        local.AppendTransformed(DeviceCloud(f), T)
    ref_local = np.concatenate([O.transform_cloud(f, T) for f, T in zip(frames, rel)], axis=0)
    assert len(local) == len(ref_local)
    assert np.array_equal(local.Download(), ref_local)
    # the same assembly in one launch (b2cloud_assemble), with an empty key frame in the middle and NaN points kept
    fr2 = [f.copy() for f in frames]
    fr2[2][::50, 0] = np.nan
    fr2.insert(3, np.zeros((0, 4), np.float32))
    rel2 = rel[:3] + [np.eye(4, dtype=np.float32)] + rel[3:]
    one = DeviceCloud(np.ones((7, 4), np.float32))            # previous contents are replaced
    one.Assemble([DeviceCloud(f) for f in fr2], rel2)
    seq = DeviceCloud()
    for f, T in zip(fr2, rel2):
        seq.AppendTransformed(DeviceCloud(f), T)
    assert len(one) == len(seq) and np.array_equal(one.Download(), seq.Download(), equal_nan=True)
    one.Assemble([], [])
    assert len(one) == 0
    vf = VoxelFilter(0.6, 0.6, 0.6)
    filt_dev = vf.FilterCloud(local)
    ok, filt_host = vf.Filter(ref_local)
    o_filt = O.voxel_filter(ref_local, 0.6, 0.6, 0.6)[0]
    assert np.array_equal(filt_dev.Download(), filt_host) and np.array_equal(filt_host, o_filt)
    vf.FilterCloud(local, local)                         # in place
    assert np.array_equal(local.Download(), o_filt)

    src = VoxelFilter(1.3, 1.3, 1.3).Filter(frames[3])[1]
    guess = (rel[3] @ synth.pose6_to_matrix(np.array([0.2, -0.15, 0.05, 0.005, -0.004, 0.01])).astype(np.float32)).astype(np.float32)
    reg_d = NDTRegistration(1.0, 0.1, 0.01, 30)
    reg_h = NDTRegistration(1.0, 0.1, 0.01, 30)
    reg_d.SetInputTargetCloud(local)
    reg_h.SetInputTarget(o_filt)
    assert reg_d.TargetInfo() == reg_h.TargetInfo()
    res_cloud = DeviceCloud()
    ok, rc, pose_d = reg_d.ScanMatchCloud(DeviceCloud(src), guess, res_cloud)
    ok, cloud_h, pose_h = reg_h.ScanMatch(src, guess)
    assert np.array_equal(pose_d, pose_h)
    assert reg_d.last_result["iterations"] == reg_h.last_result["iterations"]
    assert np.array_equal(res_cloud.Download(), cloud_h)
    assert reg_d.GetFitnessScore() == reg_h.GetFitnessScore()
    # and against the oracle
    grid = O.Grid(o_filt, 1.0)
    prm = O.params(step_size=float(np.float32(0.1)), trans_eps=float(np.float32(0.01)))
    ref = O.align(grid, prm, src, guess)
    assert reg_d.last_result["iterations"] == ref["iterations"]
    assert np.max(np.abs(pose_d[:3, 3] - ref["pose"][:3, 3])) <= 1e-3
    assert np.max(np.abs(reg_d.last_result["p"][3:] - ref["p"][3:])) <= 1e-4


def test_crop_then_target_like_matching_node(scene):
    """matching.cpp:166-183: BoxFilter(origin) on the global map -> SetInputTarget, all on the device."""
    from lidar_slam_b200.registration import BoxFilter, DeviceCloud, NDTRegistration
    gmap = scene.make_map(300000, 2.0)
    d_map = DeviceCloud(gmap)
    bf = BoxFilter([-40.0, 40.0, -40.0, 40.0, -40.0, 40.0])
    truth = scene.path_pose(30.0)
    bf.SetOrigin(truth[:3])
    d_local = bf.FilterCloud(d_map)
    local = O.box_filter(gmap, bf.GetEdge())
    assert np.array_equal(d_local.Download(), local)
    reg_d = NDTRegistration(1.0, 0.1, 0.01, 30)
    reg_h = NDTRegistration(1.0, 0.1, 0.01, 30)
    reg_d.SetInputTargetCloud(d_local)
    reg_h.SetInputTarget(local)
    Ld, Lh = reg_d.TargetLeaves(), reg_h.TargetLeaves()
    for k in ("idx", "n", "centroid", "mean", "icov"):
        assert np.array_equal(Ld[k], Lh[k], equal_nan=True), k


def test_remove_nan_keeps_finite_points_in_order(scene):
    """pcl::removeNaNFromPointCloud as front_end.cpp:92 / matching.cpp:188 call it before Filter + ScanMatch."""
    from lidar_slam_b200.registration import DeviceCloud
    scan = scene.scan(9, scene.path_pose(12.0)).copy()
    rng = np.random.default_rng(2)
    bad = rng.choice(len(scan), 500, replace=False)
    scan[bad[:200], 0] = np.nan
    scan[bad[200:350], 1] = np.inf
    scan[bad[350:], 2] = -np.inf
    scan[bad[:10], 3] = np.nan            # intensity is not tested by PCL
    out = DeviceCloud(scan).RemoveNaN().Download()
    keep = np.isfinite(scan[:, :3]).all(axis=1)
    assert len(out) == keep.sum() and np.array_equal(out, scan[keep], equal_nan=True)
    assert len(DeviceCloud(np.full((7, 4), np.nan, np.float32)).RemoveNaN()) == 0


def test_in_place_compaction_like_the_reference_calls_it(scene):
    """The reference hands its filters the same pointer as input and output (matching.cpp:158, data_pretreat_flow's
    AdjustCloud): crop, NaN removal and de-skew with dst == src give what the two-cloud call gives, repeatedly."""
    from lidar_slam_b200.registration import BoxFilter, DeviceCloud, DistortionAdjust
    scan = scene.scan(21, scene.path_pose(40.0)).copy()
    scan[::97, 1] = np.nan
    bf = BoxFilter({"box_filter_size": [-40.0, 40.0, -20.0, 20.0, -3.0, 8.0]})
    bf.SetOrigin([0.0, 0.0, 0.0])
    da = DistortionAdjust()
    da.SetMotionInfo(0.1, [8.0, 0.2, 0.0], [0.0, 0.01, 0.3])
    c = DeviceCloud(scan)
    want = c.RemoveNaN()
    assert c.RemoveNaN(c) is c and np.array_equal(c.Download(), want.Download())
    want = bf.FilterCloud(c)
    n_before = len(c)
    assert bf.FilterCloud(c, c) is c and 0 < len(c) < n_before and np.array_equal(c.Download(), want.Download())
    assert np.array_equal(bf.FilterCloud(c, c).Download(), want.Download())          # idempotent, buffers swap back
    want = da.AdjustCloudDevice(c)
    assert da.AdjustCloudDevice(c, c) is c and np.array_equal(c.Download(), want.Download())
    # the cloud is still a normal cloud afterwards: append and target build work on the swapped buffer
    d = DeviceCloud()
    d.AppendTransformed(c, np.eye(4, dtype=np.float32))
    assert np.array_equal(d.Download(), c.Download())
    e = DeviceCloud(np.zeros((0, 4), np.float32))
    assert len(bf.FilterCloud(e, e)) == 0
