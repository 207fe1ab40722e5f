"""sensor_msgs/PointCloud2 payloads and PCD files into pcl::PointXYZRGB rows.
CPU tests pin the oracle restatement (oracle/cloud_io.py) on hand-written files and round trips; GPU tests compare
gicpb_pointcloud2_to_xyzrgb / gicpb_pcd_load_xyzrgb with it bit for bit (reference call sites: src/node.cpp:37,41,
src/load_and_publish_clouds.cpp:75, src/Utils.cpp:100-105)."""
import os

import numpy as np
import pytest

from oracle import cloud_io as oio


def _cloud(n, seed=0, with_nan=False):
    rng = np.random.default_rng(seed)
    xyz = (rng.random((n, 3)) * 10 - 5).astype(np.float32)
    if with_nan and n > 5:
        xyz[3, 1] = np.nan
    rgba = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    inten = rng.random(n).astype(np.float32)
    ring = rng.integers(0, 64, n).astype(np.uint16)
    nrm = rng.normal(size=(n, 3)).astype(np.float32)
    return xyz, rgba, inten, ring, nrm


def _write_variants(tmp_path, n=257, with_nan=False):
    """PCD files of one cloud in several field layouts x the three encodings; returns [(path, xyz, rgba or None)]."""
    xyz, rgba, inten, ring, nrm = _cloud(n, 1, with_nan)
    out = []
    for kind in ("ascii", "binary", "binary_compressed"):
        # pcl::PointXYZRGB as PCDWriter writes it: x y z rgb (rgb as FLOAT32 bits in binary bodies, as rgba U4 in ascii)
        rgb_field = ("rgba", 4, "U", 1) if kind == "ascii" else ("rgb", 4, "F", 1)
        rgb_col = rgba if kind == "ascii" else rgba.view(np.float32)
        p = str(tmp_path / f"xyzrgb_{kind}.pcd")
        oio.pcd_write(p, [("x", 4, "F", 1), ("y", 4, "F", 1), ("z", 4, "F", 1), rgb_field],
                      [xyz[:, 0], xyz[:, 1], xyz[:, 2], rgb_col], kind)
        out.append((p, xyz, rgba))
        # extra fields around and between, xyz not adjacent, no colour, tabs, no comment line
        p = str(tmp_path / f"mixed_{kind}.pcd")
        oio.pcd_write(p, [("intensity", 4, "F", 1), ("x", 4, "F", 1), ("ring", 2, "U", 1), ("y", 4, "F", 1),
                          ("normal", 4, "F", 3), ("z", 4, "F", 1)],
                      [inten, xyz[:, 0], ring, xyz[:, 1], nrm, xyz[:, 2]], kind, comments=False, tabs=True)
        out.append((p, xyz, None))
        # organised cloud (HEIGHT > 1)
        if n % 2 == 0:
            p = str(tmp_path / f"organised_{kind}.pcd")
            oio.pcd_write(p, [("x", 4, "F", 1), ("y", 4, "F", 1), ("z", 4, "F", 1)], [xyz[:, 0], xyz[:, 1], xyz[:, 2]], kind,
                          width=n // 2, height=2)
            out.append((p, xyz, None))
    return out


def _expect_rows(xyz, rgba):
    rows = np.zeros((len(xyz), 8), np.float32)
    rows[:, :3] = xyz
    rows[:, 3] = 1.0
    rows.view(np.uint32)[:, 4] = 0xFF000000 if rgba is None else rgba
    return rows


def _same_bits(a, b):
    return a.shape == b.shape and np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


# ---- oracle pins (CPU) ------------------------------------------------------------------------------------
def test_lzf_round_trip_and_known_stream():
    rng = np.random.default_rng(3)
    for data in (b"", b"a", b"abcabcabcabcabcabc" * 40, rng.integers(0, 256, 5000, dtype=np.uint8).tobytes(),
                 bytes(3000), (np.arange(4000) % 7).astype(np.uint8).tobytes()):
        comp = oio.lzf_compress(data)
        assert oio.lzf_decompress(comp, len(data)) == data
    assert len(oio.lzf_compress(bytes(3000))) < 100  # back references are really used
    # hand-assembled stream: literal "ab", then a 6-byte match at distance 2, then literal "!"
    stream = bytes([1, ord("a"), ord("b"), (4 << 5) | 0, 1, 0, ord("!")])
    assert oio.lzf_decompress(stream, 9) == b"abababab!"
    # long match: length field 7 + extension byte (2 + 7 + 3 = 12 bytes at distance 1)
    assert oio.lzf_decompress(bytes([0, ord("z"), (7 << 5) | 0, 3, 0]), 13) == b"z" * 13


def test_oracle_reads_handwritten_ascii(tmp_path):
    p = tmp_path / "hand.pcd"
    p.write_text("# .PCD v.7 - Point Cloud Data file format\nVERSION .7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F F\n"
                 "COUNT 1 1 1 1\nWIDTH 3\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 3\nDATA ascii\n"
                 "0.93773 0.33763 0 4.2108e+06\n0.90805 0.35641 0 4.2108e+06\nnan 1.5 -2 4.808e+06\n")
    rows, h = oio.pcd_load_xyzrgb(str(p))
    assert h["points"] == 3 and h["point_step"] == 16 and h["data_kind"] == 0 and not h["is_dense"]
    assert rows[0, 0] == np.float32(0.93773) and rows[1, 1] == np.float32(0.35641) and np.isnan(rows[2, 0])
    assert rows[2, 2] == -2 and np.all(rows[:, 3] == 1)
    assert rows.view(np.uint32)[0, 4] == np.float32(4.2108e+06).view(np.uint32)


def test_oracle_round_trips_all_encodings(tmp_path):
    for path, xyz, rgba in _write_variants(tmp_path, 64, with_nan=True):
        rows, h = oio.pcd_load_xyzrgb(path)
        assert _same_bits(rows, _expect_rows(xyz, rgba)), path
        assert h["points"] == 64 and not h["is_dense"], path
    for path, xyz, rgba in _write_variants(tmp_path, 10):
        assert oio.pcd_read(path)["is_dense"], path


def test_oracle_header_defaults(tmp_path):
    p = tmp_path / "nohdr.pcd"  # no SIZE/TYPE/COUNT/HEIGHT/POINTS: float32 fields, one row
    p.write_text("FIELDS x y z\nWIDTH 2\nDATA ascii\n1 2 3\n4 5 6\n")
    rows, h = oio.pcd_load_xyzrgb(str(p))
    assert h["height"] == 1 and h["points"] == 2 and h["point_step"] == 12
    assert np.array_equal(rows[:, :3], np.array([[1, 2, 3], [4, 5, 6]], np.float32))
    assert np.all(rows.view(np.uint32)[:, 4] == 0xFF000000)


# ---- GPU parity ----------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def engine():
    from leica_point_cloud_processing_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.gpu
def test_pcd_load_matches_oracle(engine, tmp_path):
    for n, with_nan in ((257, False), (64, True), (1, False)):
        for path, xyz, rgba in _write_variants(tmp_path, n, with_nan):
            rows, info = engine.load_pcd(path)
            ref, h = oio.pcd_load_xyzrgb(path)
            assert _same_bits(rows, ref), path
            assert info["points"] == h["points"] and info["width"] == h["width"] and info["height"] == h["height"], path
            assert info["point_step"] == h["point_step"] and info["data_kind"] == h["data_kind"], path
            assert bool(info["is_dense"]) == h["is_dense"], path
            assert (info["off_x"], info["off_y"], info["off_z"], info["off_rgb"]) == oio.field_offsets(h["fields"]), path


@pytest.mark.gpu
def test_pcd_handwritten_padding_and_device_output(engine, tmp_path):
    import torch
    p = tmp_path / "pad.pcd"  # "_" padding field with COUNT 3 (its tokens are skipped), Windows line ends
    p.write_bytes(b"VERSION 0.7\r\nFIELDS x _ y z rgba\r\nSIZE 4 1 4 4 4\r\nTYPE F U F F U\r\nCOUNT 1 3 1 1 1\r\nWIDTH 2\r\n"
                  b"HEIGHT 1\r\nPOINTS 2\r\nDATA ascii\r\n1.25 9 9 9 2.5 -3 4278190335\r\n7 0 0 0 8 9 16711680\r\n")
    rows, info = engine.load_pcd(str(p), device_out=True)
    ref, h = oio.pcd_load_xyzrgb(str(p))
    assert rows.is_cuda and _same_bits(rows.cpu().numpy(), ref)
    assert info["point_step"] == 19 and info["off_y"] == 7  # unaligned offsets: the byte-wise gather path
    assert ref.view(np.uint32)[0, 4] == 4278190335 and ref[1, 2] == 9
    assert isinstance(rows, torch.Tensor)


@pytest.mark.gpu
def test_pcd_errors(engine, tmp_path):
    from leica_point_cloud_processing_b200._capi import GicpError
    with pytest.raises(GicpError):
        engine.load_pcd(str(tmp_path / "missing.pcd"))
    p = tmp_path / "noxyz.pcd"
    p.write_text("FIELDS a b\nWIDTH 1\nDATA ascii\n1 2\n")
    with pytest.raises(GicpError):
        engine.load_pcd(str(p))
    p = tmp_path / "short.pcd"
    p.write_bytes(b"FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 4\nHEIGHT 1\nPOINTS 4\nDATA binary\n" + bytes(20))
    with pytest.raises(GicpError):
        engine.load_pcd(str(p))
    p = tmp_path / "badlzf.pcd"
    p.write_bytes(b"FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 4\nHEIGHT 1\nPOINTS 4\nDATA binary_compressed\n" +
                  np.array([3, 48], np.uint32).tobytes() + bytes([200, 5, 1]))
    with pytest.raises(GicpError):
        engine.load_pcd(str(p))
    assert engine.pcd_info(str(p))["points"] == 4  # the header alone is fine


@pytest.mark.gpu
def test_pointcloud2_layouts_match_oracle(engine):
    import torch
    n = 1000
    xyz, rgba, inten, ring, nrm = _cloud(n, 7, with_nan=True)
    # (a) the message pcl::toROSMsg writes for PointXYZRGB: x@0 y@4 z@8 rgb@16, point_step 32
    msg = np.zeros((n, 32), np.uint8)
    msg[:, 0:12] = xyz.view(np.uint8).reshape(n, 12)
    msg[:, 12:16] = np.frombuffer(np.float32(1).tobytes(), np.uint8)
    msg[:, 16:20] = rgba.view(np.uint8).reshape(n, 4)
    rows = engine.pointcloud2_to_xyzrgb(msg, n, 1, 32, 32 * n, 0, 4, 8, 16)
    assert _same_bits(rows, _expect_rows(xyz, rgba))
    assert _same_bits(rows, oio.pc2_to_xyzrgb(msg.tobytes(), n, 1, 32, 32 * n, 0, 4, 8, 16))
    assert rows.tobytes() == msg.tobytes()  # toROSMsg is the plain copy of these rows
    # (b) scattered fields, no colour, organised with row padding (row_step > width * point_step)
    w, hgt, step = 50, 20, 28
    row_step = w * step + 12
    buf = np.random.default_rng(5).integers(0, 256, hgt * row_step, dtype=np.uint8)
    for i in range(n):
        base = (i // w) * row_step + (i % w) * step
        buf[base + 4: base + 8] = np.frombuffer(xyz[i, 0].tobytes(), np.uint8)
        buf[base + 12: base + 16] = np.frombuffer(xyz[i, 1].tobytes(), np.uint8)
        buf[base + 20: base + 24] = np.frombuffer(xyz[i, 2].tobytes(), np.uint8)
    rows = engine.pointcloud2_to_xyzrgb(buf, w, hgt, step, row_step, 4, 12, 20, -1)
    assert _same_bits(rows, _expect_rows(xyz, None))
    assert _same_bits(rows, oio.pc2_to_xyzrgb(buf.tobytes(), w, hgt, step, row_step, 4, 12, 20, -1))
    # (c) a 13-byte point (x y z + one intensity byte): unaligned gather; device in, device out
    buf = np.zeros((n, 13), np.uint8)
    buf[:, :12] = xyz.view(np.uint8).reshape(n, 12)
    buf[:, 12] = 77
    d = torch.from_numpy(buf).cuda()
    rows = engine.pointcloud2_to_xyzrgb(d, n, 1, 13, 13 * n, 0, 4, 8, -1, device_out=True)
    assert _same_bits(rows.cpu().numpy(), _expect_rows(xyz, None))
    # (d) empty message
    assert engine.pointcloud2_to_xyzrgb(b"", 0, 1, 32, 0, 0, 4, 8, 16).shape == (0, 8)


@pytest.mark.gpu
def test_pointcloud2_bad_layouts(engine):
    from leica_point_cloud_processing_b200._capi import GicpError
    buf = np.zeros(320, np.uint8)
    for args in ((10, 1, 32, 320, 0, 4, 30, -1), (10, 1, 32, 100, 0, 4, 8, -1), (10, 1, 8, 320, 0, 4, 8, -1),
                 (10, 1, 32, 320, 0, 4, 8, 29), (10, 1, 32, 320, -4, 4, 8, -1)):
        with pytest.raises(GicpError):
            engine.pointcloud2_to_xyzrgb(buf, *args)


@pytest.mark.gpu
def test_message_with_adjacent_xyz_is_read_in_place(engine):
    """A payload whose x, y, z are consecutive floats needs no unpacking: set_source(data + off_x, stride = point_step)
    answers exactly like the unpacked rows."""
    n = 5000
    xyz, rgba, inten, ring, nrm = _cloud(n, 11)
    step = 24  # intensity@0, x@4 y@8 z@12, ring@16, pad
    buf = np.zeros((n, step), np.uint8)
    buf[:, 0:4] = inten.view(np.uint8).reshape(n, 4)
    buf[:, 4:16] = xyz.view(np.uint8).reshape(n, 12)
    rows = engine.pointcloud2_to_xyzrgb(buf, n, 1, step, step * n, 4, 8, 12, -1)
    queries = (xyz[::7] + np.float32(0.01)).astype(np.float32)
    engine.set_target(rows)
    i0, d0 = engine.nn1(queries)
    in_place = np.lib.stride_tricks.as_strided(buf.reshape(-1)[4:].view(np.float32), shape=(n, 3), strides=(step, 4))
    engine.set_target(in_place)
    i1, d1 = engine.nn1(queries)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)


@pytest.mark.gpu
def test_pcd_binary_1M_points(engine, tmp_path):
    import time
    n = 1_000_000
    xyz, rgba, *_ = _cloud(n, 21)
    p = str(tmp_path / "big.pcd")
    oio.pcd_write(p, [("x", 4, "F", 1), ("y", 4, "F", 1), ("z", 4, "F", 1), ("rgb", 4, "F", 1)],
                  [xyz[:, 0], xyz[:, 1], xyz[:, 2], rgba.view(np.float32)], "binary")
    engine.load_pcd(p)
    t0 = time.perf_counter()
    rows, info = engine.load_pcd(p, device_out=True)
    dt = time.perf_counter() - t0
    print(f"PCD binary, {n} points ({os.path.getsize(p) / 1e6:.0f} MB) -> PointXYZRGB rows on the device: {dt * 1e3:.1f} ms")
    assert info["points"] == n
    assert _same_bits(rows.cpu().numpy(), _expect_rows(xyz, rgba))
