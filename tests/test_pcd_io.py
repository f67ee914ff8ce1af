"""PCD v0.7 file I/O: the on-disk format the reference reads with pcl::io::loadPCDFile
(matching.cpp:155) and writes with pcl::io::savePCDFileBinary (back_end.cpp:194).  Host-only code of libb2ndt.so."""
import struct

import numpy as np
import pytest

from lidar_slam_b200 import capi


def test_binary_roundtrip_and_header(tmp_path):
    rng = np.random.default_rng(0)
    c = rng.normal(size=(1000, 4)).astype(np.float32)
    c[3, 0] = np.nan
    p = tmp_path / "a.pcd"
    capi.pcd_write_binary(str(p), c)
    raw = p.read_bytes()
    header, payload = raw.split(b"DATA binary\n", 1)
    assert header.decode().splitlines() == ["# .PCD v0.7 - Point Cloud Data file format", "VERSION 0.7", "FIELDS x y z intensity",
                                            "SIZE 4 4 4 4", "TYPE F F F F", "COUNT 1 1 1 1", "WIDTH 1000", "HEIGHT 1",
                                            "VIEWPOINT 0 0 0 1 0 0 0", "POINTS 1000"]
    assert payload == c.tobytes()                      # PointXYZI on disk == packed {x,y,z,intensity}
    back = capi.pcd_read(str(p))
    assert np.array_equal(back, c, equal_nan=True)
    # (n,8) PointXYZI memory layout is accepted too
    c8 = np.zeros((1000, 8), np.float32); c8[:, :3] = c[:, :3]; c8[:, 3] = 1; c8[:, 4] = c[:, 3]
    capi.pcd_write_binary(str(p), c8)
    assert np.array_equal(capi.pcd_read(str(p)), c, equal_nan=True)
    capi.pcd_write_binary(str(p), np.zeros((0, 4), np.float32))
    assert capi.pcd_read(str(p)).shape == (0, 4)


def test_ascii_and_foreign_field_layouts(tmp_path):
    p = tmp_path / "b.pcd"
    p.write_text("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb intensity\nSIZE 4 4 4 4 4\nTYPE F F F U F\n"
                 "COUNT 1 1 1 1 1\nWIDTH 3\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 3\nDATA ascii\n"
                 "1.5 2 3 4278190080 0.25\nnan 0 0 0 1\n-1e3 2.5e-1 7 255 0.5\n")
    out = capi.pcd_read(str(p))
    assert np.array_equal(out, np.array([[1.5, 2, 3, 0.25], [np.nan, 0, 0, 1], [-1000, 0.25, 7, 0.5]], np.float32), equal_nan=True)
    # binary with intensity first, a double z, a 3-count normal and no padding; xyz only -> intensity 0
    q = tmp_path / "c.pcd"
    hdr = ("VERSION 0.7\nFIELDS intensity x y z normal\nSIZE 4 4 4 8 4\nTYPE F F F F F\nCOUNT 1 1 1 1 3\nWIDTH 2\nHEIGHT 1\n"
           "POINTS 2\nDATA binary\n").encode()
    rec = struct.pack("<fffdfff", 0.5, 1, 2, 3.0, 9, 9, 9) + struct.pack("<fffdfff", 0.75, -1, -2, -3.0, 8, 8, 8)
    q.write_bytes(hdr + rec)
    assert np.array_equal(capi.pcd_read(str(q)), np.array([[1, 2, 3, 0.5], [-1, -2, -3, 0.75]], np.float32))
    r = tmp_path / "d.pcd"
    r.write_bytes(b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary\n" + struct.pack("<fff", 4, 5, 6))
    assert np.array_equal(capi.pcd_read(str(r)), np.array([[4, 5, 6, 0]], np.float32))


def test_errors_are_reported(tmp_path):
    with pytest.raises(capi.B2Error) as e:
        capi.pcd_read(str(tmp_path / "missing.pcd"))
    assert "cannot open" in str(e.value)
    t = tmp_path / "t.pcd"
    t.write_bytes(b"VERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 5\nHEIGHT 1\nPOINTS 5\nDATA binary\n" + b"\0" * 40)
    with pytest.raises(capi.B2Error) as e:
        capi.pcd_read(str(t))
    assert "truncated" in str(e.value)
    z = tmp_path / "z.pcd"
    z.write_bytes(b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary_compressed\n")
    with pytest.raises(capi.B2Error) as e:
        capi.pcd_read(str(z))
    assert "not supported" in str(e.value)
    n = tmp_path / "n.pcd"
    n.write_text("hello\n")
    with pytest.raises(capi.B2Error):
        capi.pcd_read(str(n))


@pytest.mark.gpu
def test_device_cloud_load_save(tmp_path):
    from lidar_slam_b200.registration import DeviceCloud
    rng = np.random.default_rng(1)
    c = rng.uniform(-50, 50, size=(20000, 4)).astype(np.float32)
    p = tmp_path / "m.pcd"
    DeviceCloud(c).SavePCD(str(p))
    assert np.array_equal(capi.pcd_read(str(p)), c)
    d = DeviceCloud().LoadPCD(str(p))
    assert len(d) == len(c) and np.array_equal(d.Download(), c)
