"""Scan de-skew (DistortionAdjust, data_pretreat): the oracle's restatement is pinned to outputs of
the reference's OWN distortion_adjust.cpp (tests/golden/deskew.npz, made by tools/make_golden.py from
oracle/_ref/libdeskew_ref.so); the GPU path is checked against both.  The reference computes in float through Eigen,
the oracle in float64: agreement is to float round-off of ~100 m coordinates."""
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deskew.npz")
TOL = 2e-4      # metres


def cases():
    g = np.load(GOLD)
    for i in range(len(g["period"])):
        yield g["scan"], g["lin"][i], g["ang"][i], float(g["period"][i]), g["out%d" % i]


def test_oracle_matches_reference_outputs():
    for scan, lin, ang, period, ref in cases():
        out = O.distortion_adjust(scan, period, lin, ang)
        assert out.shape == ref.shape
        assert np.max(np.abs(out - ref)) <= TOL
        assert np.all(out[:, 3] == 0) and np.all(ref[:, 3] == 0)         # intensity is not carried over
    # zero motion: the kept points come back where they were (to float round-off), point 0 and the 5 degree sector are gone
    scan, lin, ang, period, ref = list(cases())[1]
    az = np.arctan2(scan[:, 1], scan[:, 0]) - np.arctan2(scan[0, 1], scan[0, 0])
    az = np.mod(az, 2 * np.pi)
    keep = ~((az < np.radians(5.0)) | (2 * np.pi - az < np.radians(5.0)))
    keep[0] = False
    assert keep.sum() == len(ref) and np.max(np.abs(ref[:, :3] - scan[keep, :3])) <= TOL


def test_oracle_against_reference_library_when_built():
    if O.deskew_ref_lib() is None:
        pytest.skip("oracle/_ref/libdeskew_ref.so not built (reference tree absent)")
    rng = np.random.default_rng(3)
    scan = np.concatenate([rng.uniform(-60, 60, (5000, 2)), rng.uniform(-2, 8, (5000, 1)), rng.uniform(0, 1, (5000, 1))], 1).astype(np.float32)
    lin, ang = [5.0, -1.0, 0.2], [0.1, -0.05, 0.5]
    ref = O.distortion_adjust_reference(scan, 0.1, lin, ang)
    out = O.distortion_adjust(scan, 0.1, lin, ang)
    assert out.shape == ref.shape and np.max(np.abs(out - ref)) <= TOL


@pytest.mark.gpu
def test_gpu_deskew_matches_reference_and_oracle():
    from lidar_slam_b200.registration import DeviceCloud, DistortionAdjust
    da = DistortionAdjust()
    for scan, lin, ang, period, ref in cases():
        da.SetMotionInfo(period, lin, ang)
        ok, out = da.AdjustCloud(scan)
        assert ok and out.shape == ref.shape
        assert np.max(np.abs(out - ref)) <= TOL and np.all(out[:, 3] == 0)
        assert np.max(np.abs(out - O.distortion_adjust(scan, period, lin, ang))) <= TOL
    # device-resident, empty and single-point clouds
    assert len(da.AdjustCloudDevice(DeviceCloud(np.zeros((0, 4), np.float32)))) == 0
    assert len(da.AdjustCloudDevice(DeviceCloud(np.array([[1, 2, 3, 4]], np.float32)))) == 0


@pytest.mark.gpu
def test_gpu_fused_ingest_equals_the_three_call_sequence():
    """b2vf_ingest_filter_cloud (de-skew + NaN removal folded into the first kernel of the voxel filter, one pipeline
    run) against DistortionAdjust -> removeNaN -> VoxelFilter as three calls: filtered clouds bit-identical, and the
    ingested cloud with its NaN holes compacted away == the sequence's cloud.  Also without de-skew (NaN removal only),
    for a cloud uploaded from the host and for one produced on the device."""
    from lidar_slam_b200 import capi
    from lidar_slam_b200.registration import DeviceCloud, DistortionAdjust, VoxelFilter
    da = DistortionAdjust()
    vf = VoxelFilter(1.3, 1.3, 1.3)
    for scan, lin, ang, period, ref in cases():
        scan = scan.copy()
        scan[5::97, 1] = np.nan                       # the reference removes NaNs after the de-skew (front_end.cpp:92)
        scan[11::301, 2] = np.inf
        for on_device in (False, True):
            src = DeviceCloud(scan)
            if on_device:                              # a cloud written by a device operation: no cached first point
                tmp = DeviceCloud()
                tmp.AppendTransformed(src, np.eye(4, dtype=np.float32))
                src = tmp
            da.SetMotionInfo(period, lin, ang)
            seq = da.AdjustCloudDevice(src).RemoveNaN()
            seq_f = vf.FilterCloud(seq)
            l0 = capi.launches()
            ing = DeviceCloud()
            fused_f, _ = vf.IngestFilterCloud(src, ingested=ing, scan_period=period, linear_velocity=lin, angular_velocity=ang)
            fused_launches = capi.launches() - l0
            assert np.array_equal(fused_f.Download(), seq_f.Download())
            assert len(ing) == len(scan)
            holes = ing.Download()
            kept = holes[np.isfinite(holes[:, :3]).all(axis=1)]
            assert np.array_equal(kept, seq.Download()) and np.array_equal(ing.RemoveNaN().Download(), seq.Download())
            assert fused_launches <= 10, fused_launches   # the whole ingest + filter: bbox/ingest, keys, <= 4 passes, runs, 2 centroid kernels
    # NaN removal only
    src = DeviceCloud(scan)
    a = vf.FilterCloud(src.RemoveNaN()).Download()
    b, _ = vf.IngestFilterCloud(src)
    assert np.array_equal(b.Download(), a)
