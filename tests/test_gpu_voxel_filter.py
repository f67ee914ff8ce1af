"""-m gpu: VoxelFilter (CUDA radix-sort-by-voxel-key + segmented reduce) through the C ABI vs the oracle.
Bar: voxel ids and point counts bit-exact; centroids bit-exact for the oracle's (input-order) member
order, and within few-ulp*n of any other valid PCL order (PCL's std::sort is unstable, Appendix A.3)."""
import os

import numpy as np
import pytest

from lidar_slam_b200.registration import VoxelFilter, to_xyzi8

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ndt_small.npz")


def _check(oracle, cloud, leaf):
    vf = VoxelFilter(*leaf)
    ok, out, idx, cnt = vf.Filter(cloud, with_info=True)
    o_out, o_idx, o_cnt, ov = oracle.voxel_filter(cloud, *leaf)
    assert ok
    assert np.array_equal(idx, o_idx), "voxel ids differ"
    assert np.array_equal(cnt, o_cnt), "point counts differ"
    if cloud.shape[1] == 8:
        assert np.array_equal(out[:, :3], o_out[:, :3]) and np.array_equal(out[:, 4], o_out[:, 3])
        assert np.all(out[:, 3] == 1.0) and np.all(out[:, 5:] == 0.0)
    else:
        assert np.array_equal(out, o_out), "centroids differ"
    return out, idx, cnt


def test_scan_filters_bit_exact(oracle, scans):
    for (_, scan), leaf in zip(scans, [(1.3, 1.3, 1.3), (0.6, 0.6, 0.6), (0.5, 0.5, 0.5), (1.5, 1.0, 0.7)]):
        _check(oracle, scan, leaf)


def test_golden_fixture(oracle):
    G = np.load(GOLD)
    vf = VoxelFilter(1.3, 1.3, 1.3)
    ok, out, idx, cnt = vf.Filter(G["raw"], with_info=True)
    assert np.array_equal(idx, G["raw_idx"]) and np.array_equal(cnt, G["raw_cnt"]) and np.array_equal(out, G["raw_filt"])


def test_pointxyzi_layout_and_inplace(oracle, scans):
    _, scan = scans[0]
    c8 = to_xyzi8(scan)
    out8, idx, cnt = _check(oracle, c8, (1.3, 1.3, 1.3))
    # in == out through the raw ABI (matching.cpp:158, viewer.cpp:207)
    import ctypes as C
    from lidar_slam_b200 import capi
    buf = c8.copy()
    vf = VoxelFilter(1.3, 1.3, 1.3)
    m = C.c_size_t(0)
    capi.check(capi.lib().b2vf_filter(vf._h, buf.ctypes.data, len(buf), 32, 16, buf.ctypes.data, len(buf), 32, 16, C.byref(m), None, None))
    assert m.value == len(out8) and np.array_equal(buf[:m.value], out8)


def test_map_sized_cloud(oracle, small_map):
    _check(oracle, small_map, (0.6, 0.6, 0.6))
    _check(oracle, small_map, (0.3, 0.3, 0.3))


def test_edge_cases(oracle):
    vf = VoxelFilter(1.0, 1.0, 1.0)
    ok, out = vf.Filter(np.zeros((0, 4), np.float32))
    assert ok and len(out) == 0
    pts = np.array([[0.5, 0.5, 0.5, 1], [np.nan, 0, 0, 2], [0.6, 0.6, np.inf, 3], [0.7, 0.5, 0.5, 5]], np.float32)
    _check(oracle, pts, (1.0, 1.0, 1.0))
    ok, out = vf.Filter(np.full((5, 4), np.nan, np.float32))
    assert len(out) == 0
    _check(oracle, np.array([[3, 4, 5, 9]], np.float32), (0.5, 0.5, 0.5))
    # PCL's int32 guard: output = input
    far = np.array([[0, 0, 0, 1], [5000, 5000, 5000, 2]], np.float32)
    ok, out = VoxelFilter(0.01, 0.01, 0.01).Filter(far)
    assert np.array_equal(out, far)
    # ragged sizes around the 4096-key tile and 32-lane boundaries
    rng = np.random.default_rng(1)
    for n in (1, 31, 32, 33, 4095, 4096, 4097, 8193, 70001):
        c = rng.uniform(-30, 30, (n, 4)).astype(np.float32)
        _check(oracle, c, (0.9, 1.1, 1.3))
    # negative coordinates + duplicates + everything in one voxel
    c = np.tile(np.array([[-3.2, -7.7, -0.1, 0.5]], np.float32), (1000, 1))
    out, idx, cnt = _check(oracle, c, (1.0, 1.0, 1.0))
    assert len(out) == 1 and cnt[0] == 1000


def test_full_size_properties(oracle):
    """BASELINE-size inputs (5 M-point map, 0.6 m leaf): size-independent properties + oracle spot check."""
    rng = np.random.default_rng(2)
    n = 5_000_000
    c = np.empty((n, 4), np.float32)
    c[:, 0] = rng.uniform(-100, 1100, n); c[:, 1] = rng.uniform(-150, 150, n); c[:, 2] = rng.normal(0, 3, n); c[:, 3] = rng.uniform(0, 1, n)
    vf = VoxelFilter(0.6, 0.6, 0.6)
    ok, out, idx, cnt = vf.Filter(c, with_info=True)
    assert cnt.sum() == n and np.all(np.diff(idx) > 0) and cnt.min() >= 1
    # count-weighted centroid mean == cloud mean
    w = (out[:, :3].astype(np.float64) * cnt[:, None]).sum(0) / n
    assert np.allclose(w, c[:, :3].astype(np.float64).mean(0), atol=1e-3)
    # every output point lies in its voxel
    lay = oracle.vox_layout(c, 0.6, 0.6, 0.6)
    inv = np.float32(1.0) / np.float32(0.6)
    ijk = np.floor(out[:, :3] * inv).astype(np.int64) - np.array(lay.min_b[:])
    lin = ijk[:, 0] + ijk[:, 1] * lay.div_b[0] + ijk[:, 2] * lay.div_b[0] * lay.div_b[1]
    assert np.mean(lin == idx) > 0.9999
    o_out, o_idx, o_cnt, _ = oracle.voxel_filter(c, 0.6, 0.6, 0.6)
    assert np.array_equal(idx, o_idx) and np.array_equal(cnt, o_cnt) and np.array_equal(out, o_out)
    # filtering again keeps one point per voxel
    ok, out2, idx2, cnt2 = vf.Filter(out, with_info=True)
    assert cnt2.sum() == len(out)
