"""Initial-yaw search of the matching node: Gaussian height grid + 270-bin yaw scan on the
device, against the oracle's restatement of Matching::generateGauss2DMapCells / getInitialYawAngle."""
import numpy as np
import pytest

from lidar_slam_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    scene = synth.Scene(leg=60.0)
    gmap = scene.make_map(120000, 2.0)
    truth = scene.path_pose(25.0)
    edge = [truth[0] - 40, truth[0] + 40, truth[1] - 40, truth[1] + 40, truth[2] - 40, truth[2] + 40]
    local = O.box_filter(gmap, edge)[:30000].copy()           # the oracle's build is a Python loop: keep it small
    local[5, 0] = np.nan
    return scene, local, truth


def test_height_grid_matches_reference_recurrence_bit_for_bit(world):
    from lidar_slam_b200.registration import DeviceCloud, InitialYawSearch
    scene, local, truth = world
    ys = InitialYawSearch(0.8)
    info = ys.GenerateGauss2DMapCells(DeviceCloud(local), truth[:3])
    ref = O.gauss2d_map_cells(local, truth[:3].astype(np.float32), 0.8)
    assert info["width"] == ref["width"] and info["height"] == ref["height"]
    assert np.array_equal(info["min_xyz"], ref["min_xyz"]) and np.array_equal(info["max_xyz"], ref["max_xyz"])
    mu, sigma, cnt = ys.Cells()
    assert np.array_equal(cnt, ref["cnt"])
    assert np.array_equal(mu, ref["mu"]) and np.array_equal(sigma, ref["sigma"])     # same float ops in the same order
    assert cnt.max() > 20 and (cnt == 1).sum() > 0                                   # crowded and single-point cells


def test_yaw_search_finds_heading_and_agrees_with_oracle(world):
    from lidar_slam_b200.registration import DeviceCloud, InitialYawSearch
    scene, local, truth = world
    ys = InitialYawSearch(0.8)
    ys.GenerateGauss2DMapCells(DeviceCloud(local), truth[:3])
    ref_cells = O.gauss2d_map_cells(local, truth[:3].astype(np.float32), 0.8)
    scan = scene.scan(4242, truth)[::8].copy()                 # sensor frame
    scan[3, 1] = np.inf
    best, probs = ys.GetInitialYawAngle(DeviceCloud(scan), 270)
    o_best, o_probs = O.initial_yaw_angle(ref_cells, scan, 270)
    fin = np.isfinite(o_probs)
    assert np.array_equal(np.isfinite(probs), fin)             # NaN bins (z == mu in a single-point cell) reproduce
    assert np.allclose(probs[fin], o_probs[fin], rtol=1e-6, atol=1e-9)
    assert best == o_best
    # the winning bin is the vehicle's heading (yaw of the truth pose), to the bin width
    yaw = float(truth[5]) % (2 * np.pi)
    d = abs((best - yaw + np.pi) % (2 * np.pi) - np.pi)
    assert d <= 2 * (2 * np.pi / 270) + 0.02, (best, yaw)
    # empty scan / empty map
    assert ys.GetInitialYawAngle(DeviceCloud(np.zeros((0, 4), np.float32)), 270)[0] == 0.0
    ys2 = InitialYawSearch(0.8)
    info = ys2.GenerateGauss2DMapCells(DeviceCloud(np.zeros((0, 4), np.float32)), [0, 0, 0])
    assert info["width"] == 0 and ys2.GetInitialYawAngle(DeviceCloud(scan), 90)[0] == 0.0


def test_relocalize_from_position_prior(world):
    """Config 5 in miniature: position prior off by ~3 m, heading unknown -> yaw scan + lattice of batched NDT matches."""
    from lidar_slam_b200 import batch
    from lidar_slam_b200.registration import DeviceCloud, InitialYawSearch, NDTRegistration, VoxelFilter
    scene, local, truth = world
    gmap = scene.make_map(200000, 2.0)
    reg = NDTRegistration(1.0, 0.1, 0.01, 30)
    reg.SetInputTarget(gmap)
    prior = truth[:3] + np.array([2.5, -1.5, 0.0])
    ys = InitialYawSearch(0.8)
    edge = [prior[0] - 40, prior[0] + 40, prior[1] - 40, prior[1] + 40, prior[2] - 40, prior[2] + 40]
    ys.GenerateGauss2DMapCells(DeviceCloud(O.box_filter(gmap, edge)), prior)
    scan = scene.scan(777, truth)
    filt = VoxelFilter(1.3, 1.3, 1.3).Filter(scan)[1]
    d_scan = DeviceCloud(scan)
    # PoseSearch with a zero offset is the yaw scan
    assert np.array_equal(ys.PoseSearch(d_scan, [[0.0, 0.0]], 270)[0], ys.GetInitialYawAngle(d_scan, 270)[1], equal_nan=True)
    pose, k, poses, res = batch.relocalize(reg, ys, filt, d_scan, prior, lattice=9, pitch=1.0, top=16)
    T = synth.pose6_to_matrix(truth)
    assert len(poses) == 16
    assert np.linalg.norm(pose[:3, 3] - T[:3, 3]) < 0.15, (pose[:3, 3], T[:3, 3])
    assert np.max(np.abs(pose[:3, :3] - T[:3, :3])) < 0.02
