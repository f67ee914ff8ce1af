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
    z.write_bytes(b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary_zstd\n")
    with pytest.raises(capi.B2Error) as e:
        capi.pcd_read(str(z))
    assert "not supported" in str(e.value)
    n = tmp_path / "n.pcd"
    n.write_text("hello\n")
    with pytest.raises(capi.B2Error):
        capi.pcd_read(str(n))


def _lzf_compress(data: bytes) -> bytes:
    """Greedy LZF encoder (test helper): literal runs of <= 32 bytes, back references of 3..264 bytes within 8 KiB."""
    out, lit, i, n, last = bytearray(), bytearray(), 0, len(data), {}

    def flush():
        for k in range(0, len(lit), 32):
            run = lit[k:k + 32]
            out.append(len(run) - 1); out.extend(run)
        lit.clear()

    while i < n:
        key = data[i:i + 3]
        j = last.get(key, -1) if len(key) == 3 else -1
        last[key] = i
        if j >= 0 and i - j <= 8192:
            m = 3
            while i + m < n and m < 264 and data[j + m] == data[i + m]:
                m += 1
            flush()
            dist, ln = i - j - 1, m - 2
            if ln < 7:
                out.append((ln << 5) | (dist >> 8))
            else:
                out.append((7 << 5) | (dist >> 8)); out.append(ln - 7)
            out.append(dist & 0xFF)
            i += m
        else:
            lit.append(data[i]); i += 1
    flush()
    return bytes(out)


def test_binary_compressed(tmp_path):
    """DATA binary_compressed: u32 sizes + one LZF stream holding the cloud field by field (PCD v0.7)."""
    rng = np.random.default_rng(7)
    n = 5000
    c = np.zeros((n, 4), np.float32)
    c[:, 0] = np.round(rng.uniform(-50, 50, n), 1); c[:, 1] = np.round(rng.uniform(-50, 50, n), 1)
    c[:, 2] = 0.0                                     # a constant field: long overlapping back references
    c[:, 3] = rng.integers(0, 4, n) / 4.0
    c[17, 0] = np.nan
    # fields in a foreign order with an extra u16 ring field in the middle
    ring = rng.integers(0, 64, n).astype(np.uint16)
    raw = c[:, 3].tobytes() + c[:, 0].tobytes() + ring.tobytes() + c[:, 1].tobytes() + c[:, 2].tobytes()
    comp = _lzf_compress(raw)
    assert len(comp) < 0.7 * len(raw)                 # the helper really emits back references
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS intensity x ring y z\nSIZE 4 4 2 4 4\nTYPE F F U F F\n"
           "COUNT 1 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary_compressed\n" % (n, n)).encode()
    p = tmp_path / "bc.pcd"
    p.write_bytes(hdr + struct.pack("<II", len(comp), len(raw)) + comp)
    assert np.array_equal(capi.pcd_read(str(p)), c, equal_nan=True)
    # literal-only stream (incompressible data), empty cloud
    lit = b"".join(bytes([len(raw[k:k + 32]) - 1]) + raw[k:k + 32] for k in range(0, len(raw), 32))
    p.write_bytes(hdr + struct.pack("<II", len(lit), len(raw)) + lit)
    assert np.array_equal(capi.pcd_read(str(p)), c, equal_nan=True)
    e0 = tmp_path / "e0.pcd"
    e0.write_bytes(b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 0\nHEIGHT 1\nPOINTS 0\nDATA binary_compressed\n" + struct.pack("<II", 0, 0))
    assert capi.pcd_read(str(e0)).shape == (0, 4)
    # corrupt streams are reported, not read out of bounds: truncated payload, wrong size word, reference before the start
    for bad in (hdr + struct.pack("<II", len(comp), len(raw)) + comp[:len(comp) // 2],
                hdr + struct.pack("<II", len(comp), len(raw) - 4) + comp,
                hdr + struct.pack("<II", 3, len(raw)) + bytes([0x20, 0x05, 0x00])):
        p.write_bytes(bad)
        with pytest.raises(capi.B2Error):
            capi.pcd_read(str(p))


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


def test_compressed_reader_survives_corruption(tmp_path):
    """Random byte flips / truncations of a valid binary_compressed file either decode to n points or raise B2Error;
    the decoder never reads or writes outside its buffers (a crash would take the test process down)."""
    rng = np.random.default_rng(11)
    n = 600
    c = np.round(rng.uniform(-20, 20, (n, 4)), 1).astype(np.float32)
    raw = b"".join(c[:, k].tobytes() for k in range(4))
    comp = _lzf_compress(raw)
    hdr = ("VERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nPOINTS %d\n"
           "DATA binary_compressed\n" % (n, n)).encode()
    good = hdr + struct.pack("<II", len(comp), len(raw)) + comp
    p = tmp_path / "f.pcd"
    p.write_bytes(good)
    assert np.array_equal(capi.pcd_read(str(p)), c)
    ok = bad = 0
    for trial in range(300):
        b = bytearray(good)
        kind = trial % 3
        if kind == 0:                                   # flip bytes in the LZF payload
            for pos in rng.integers(len(hdr) + 8, len(b), rng.integers(1, 6)):
                b[pos] = rng.integers(0, 256)
        elif kind == 1:                                 # lie in the size words
            struct.pack_into("<I", b, len(hdr) + 4 * int(rng.integers(0, 2)), int(rng.integers(0, 2 * len(raw))))
        else:                                           # truncate
            b = b[:int(rng.integers(len(hdr), len(b)))]
        p.write_bytes(bytes(b))
        try:
            out = capi.pcd_read(str(p))
            assert out.shape == (n, 4)
            ok += 1
        except capi.B2Error:
            bad += 1
    assert bad > 100 and ok + bad == 300


def test_hostile_headers_are_errors_not_crashes(tmp_path):
    """Headers that ask for absurd allocations (huge COUNT / POINTS / compressed sizes) come back as B2Error."""
    p = tmp_path / "h.pcd"
    cases = [
        b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 2000000000\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA binary\n" + b"\0" * 64,
        b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 4000000000\nHEIGHT 1\nPOINTS 4000000000\nDATA binary\n",
        b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1000000000000\nHEIGHT 1000000\nDATA ascii\n1 2 3\n",
        b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA binary_compressed\n"
        + struct.pack("<II", 0xFFFFFFF0, 24) + b"\x00" * 16,
        b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 3\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary\n" + b"\0" * 16,
        # TYPE F with SIZE 1 / 2 would make the reader fetch 8 bytes from a 2-byte field (past the record)
        b"VERSION 0.7\nFIELDS x y z\nSIZE 2 2 2\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary\n" + b"\0" * 6,
        b"VERSION 0.7\nFIELDS x y z\nSIZE 1 1 1\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary\n" + b"\0" * 3,
        b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F Q F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary\n" + b"\0" * 12,
    ]
    for blob in cases:
        p.write_bytes(blob)
        with pytest.raises(capi.B2Error):
            capi.pcd_read(str(p))


def test_direct_path_needs_four_float32_fields_and_long_ascii_rows(tmp_path):
    """SIZE 8 4 2 2 also has a 16-byte record but is not x,y,z,intensity float32: it must take the per-field path;
    an ascii row longer than the reader's line buffer stays ONE point."""
    p = tmp_path / "m.pcd"
    rec = struct.pack("<dfhH", 1.5, 2.5, -3, 7)
    p.write_bytes(b"VERSION 0.7\nFIELDS x y z intensity\nSIZE 8 4 2 2\nTYPE F F I U\nCOUNT 1 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA binary\n" + rec * 2)
    assert np.array_equal(capi.pcd_read(str(p)), np.array([[1.5, 2.5, -3.0, 7.0]] * 2, np.float32))
    pad = " ".join(["0"] * 3000)
    p.write_bytes(("VERSION 0.7\nFIELDS x y z pad\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 3000\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA ascii\n"
                   "1 2 3 %s\n4 5 6 %s\n" % (pad, pad)).encode())
    assert np.array_equal(capi.pcd_read(str(p))[:, :3], np.array([[1, 2, 3], [4, 5, 6]], np.float32))
