"""CPU restatement of the wire / on-disk formats at the boundary of the registration path.

TEST INFRASTRUCTURE ONLY - the checker of tests/ and nothing else; the product path (libgicp_b200.so,
gicpb_pointcloud2_to_xyzrgb / gicpb_pcd_load_xyzrgb) never imports or calls this module.

What it follows (PCL 1.8.1 is a third-party dependency that is not vendored in /root/reference; pinned by the
reference's .travis.yml:11 / README.md:82-89; restated here from its published sources):
  * pcl::fromROSMsg -> pcl::fromPCLPointCloud2 with a MsgFieldMap (common/include/pcl/conversions.h): every field of
    the point type that the message has under the same name, datatype and count is copied, everything else keeps the
    default-constructed value of pcl::PointXYZRGB (x = y = z = 0, data[3] = 1, r = g = b = 0, a = 255).  "rgb" FLOAT32
    and "rgba" UINT32 match each other (FieldMatches specialisation).  Reference call sites: src/node.cpp:37,41.
  * pcl::PCDReader::readHeader / read (io/src/pcd_io.cpp): FIELDS / SIZE / TYPE / COUNT / WIDTH / HEIGHT / VIEWPOINT /
    POINTS / DATA; ascii bodies parsed token by token ("nan" -> NaN and is_dense = false; "_" padding fields skipped),
    binary bodies copied, binary_compressed bodies = two uint32 sizes + an LZF stream of the field-major copy
    ("xxyyzz"), then every FLOAT32 / FLOAT64 value of a binary body is tested for finiteness (is_dense).
    Reference call site: pcl::io::loadPCDFile<pcl::PointXYZRGB>, src/load_and_publish_clouds.cpp:75.
  * LZF: liblzf 3.x lzf_decompress as bundled in PCL (io/src/lzf.cpp).  The compressor below is only used to WRITE test
    files; any valid LZF stream decodes to the same bytes.

Parity: unpinned by the reference (its tests hold no PointCloud2 or PCD golden files); pinned here by round trips
through the three encodings and by hand-written files in tests/test_cloud_io.py.
"""
import numpy as np

_NP = {("F", 4): np.float32, ("F", 8): np.float64, ("I", 1): np.int8, ("I", 2): np.int16, ("I", 4): np.int32,
       ("I", 8): np.int64, ("U", 1): np.uint8, ("U", 2): np.uint16, ("U", 4): np.uint32, ("U", 8): np.uint64}


def pc2_to_xyzrgb(data, width, height, point_step, row_step, off_x, off_y, off_z, off_rgb=-1):
    """fromPCLPointCloud2 into PointXYZRGB rows: float32 [n, 8] = x, y, z, 1.0, rgba bits, 0, 0, 0."""
    buf = np.frombuffer(bytes(data), np.uint8)
    n = width * height
    out = np.zeros((n, 8), np.float32)
    out[:, 3] = 1.0
    bits = out.view(np.uint32)
    bits[:, 4] = 0xFF000000
    for i in range(n):
        row, col = divmod(i, width) if width else (0, 0)
        base = row * row_step + col * point_step
        for k, off in enumerate((off_x, off_y, off_z)):
            bits[i, k] = int.from_bytes(buf[base + off: base + off + 4].tobytes(), "little")
        if off_rgb >= 0:
            bits[i, 4] = int.from_bytes(buf[base + off_rgb: base + off_rgb + 4].tobytes(), "little")
    return out


def lzf_decompress(src, out_len):
    src = bytes(src)
    out = bytearray()
    ip = 0
    while ip < len(src):
        ctrl = src[ip]
        ip += 1
        if ctrl < 32:
            out += src[ip: ip + ctrl + 1]
            ip += ctrl + 1
        else:
            ln = ctrl >> 5
            if ln == 7:
                ln += src[ip]
                ip += 1
            dist = ((ctrl & 0x1F) << 8) + src[ip] + 1
            ip += 1
            for _ in range(ln + 2):
                out.append(out[-dist])
    assert len(out) == out_len, (len(out), out_len)
    return bytes(out)


def lzf_compress(src):
    """A small greedy LZF encoder (hash of 3-byte prefixes): literal runs <= 32, matches 3..264 bytes, distance <= 8192."""
    src = bytes(src)
    n = len(src)
    out = bytearray()
    lit = bytearray()
    table = {}

    def flush():
        nonlocal lit
        for k in range(0, len(lit), 32):
            chunk = lit[k: k + 32]
            out.append(len(chunk) - 1)
            out.extend(chunk)
        lit = bytearray()

    i = 0
    while i < n:
        key = src[i: i + 3]
        ref = table.get(key) if i + 2 < n else None
        if i + 2 < n:
            table[key] = i
        if ref is not None and 0 < i - ref <= 8192:
            ln = 3
            while i + ln < n and ln < 264 and src[ref + ln] == src[i + ln]:
                ln += 1
            flush()
            dist = i - ref - 1
            l2 = ln - 2
            if l2 < 7:
                out.append((l2 << 5) | (dist >> 8))
            else:
                out.append((7 << 5) | (dist >> 8))
                out.append(l2 - 7)
            out.append(dist & 0xFF)
            i += ln
        else:
            lit.append(src[i])
            i += 1
    flush()
    return bytes(out)


def _split(line):
    return [t for t in line.replace("\t", " ").replace("\r", " ").split(" ") if t]


def pcd_read(path):
    """PCDReader::read: returns dict(fields=[(name, offset, size, type, count)], width, height, points, point_step,
    data_kind, is_dense, blob=bytes of points * point_step)."""
    raw = open(path, "rb").read()
    pos = 0
    fields = []
    width = height = points = None
    kind = None
    step = 0
    while pos < len(raw):
        end = raw.find(b"\n", pos)
        end = len(raw) if end < 0 else end
        line = raw[pos:end].decode("ascii", "replace")
        pos = end + 1
        st = _split(line)
        if not st or st[0].startswith("#"):
            continue
        key = st[0]
        if key in ("VERSION", "VIEWPOINT"):
            continue
        if key in ("FIELDS", "COLUMNS"):
            fields = [[name, 4 * i, 4, "F", 1] for i, name in enumerate(st[1:])]
            step = 4 * len(fields)
        elif key in ("SIZE", "TYPE", "COUNT"):
            assert len(st) - 1 == len(fields)
            off = 0
            for f, tok in zip(fields, st[1:]):
                if key == "SIZE":
                    f[2] = int(tok)
                elif key == "TYPE":
                    f[3] = tok[0]
                else:
                    f[4] = int(tok)
                f[1] = off
                off += f[2] * f[4]
            step = off
        elif key == "WIDTH":
            width = int(st[1])
        elif key == "HEIGHT":
            height = int(st[1])
        elif key == "POINTS":
            points = int(st[1])
        elif key == "DATA":
            kind = {"ascii": 0, "binary": 1, "binary_compressed": 2}[st[1]]
            break
        else:
            raise ValueError("unknown PCD header entry " + key)
    if height is None:
        height = 1
        if not width and points is not None:
            width = points
    if points is None:
        points = width * height
    assert width * height == points
    total = points * step
    dense = True
    if kind == 0:
        blob = bytearray(total)
        idx = 0
        for line in raw[pos:].decode("ascii", "replace").split("\n"):
            if idx >= points:
                break
            st = _split(line)
            if not st:
                continue
            tok = 0
            for name, off, size, typ, count in fields:
                if name == "_":
                    tok += count
                    continue
                for c in range(count):
                    if tok < len(st):
                        t = st[tok]
                        dt = _NP[(typ, size)]
                        if t == "nan":
                            dense = False
                            v = np.array([np.nan if typ == "F" else 0]).astype(dt)
                        elif typ == "F":
                            v = np.array([float(t)], np.float64).astype(dt)
                        else:
                            v = np.array([int(t)]).astype(dt)
                        blob[idx * step + off + c * size: idx * step + off + (c + 1) * size] = v.tobytes()
                    tok += 1
            idx += 1
        assert idx == points
        blob = bytes(blob)
    elif kind == 1:
        blob = raw[pos: pos + total]
        assert len(blob) == total
    else:
        comp, uncomp = np.frombuffer(raw[pos: pos + 8], np.uint32)
        assert uncomp == total
        soa = lzf_decompress(raw[pos + 8: pos + 8 + int(comp)], total)
        out = np.zeros((points, step), np.uint8)
        toff = 0
        for name, off, size, typ, count in fields:
            fs = size * count
            out[:, off: off + fs] = np.frombuffer(soa[toff: toff + fs * points], np.uint8).reshape(points, fs)
            toff += fs * points
        blob = out.tobytes()
    if kind != 0:
        arr = np.frombuffer(blob, np.uint8).reshape(points, step) if points else np.zeros((0, step), np.uint8)
        for name, off, size, typ, count in fields:
            if name == "_" or typ != "F":
                continue
            vals = np.ascontiguousarray(arr[:, off: off + size * count]).view(_NP[(typ, size)])
            if not np.isfinite(vals).all():
                dense = False
    return dict(fields=[tuple(f) for f in fields], width=width, height=height, points=points, point_step=step,
                data_kind=kind, is_dense=dense, blob=blob)


def field_offsets(fields):
    """pcl::FieldMatches for PointXYZRGB: offsets of x, y, z (FLOAT32, count 1) and rgb (FLOAT32) / rgba (UINT32), -1 = absent."""
    off = {"x": -1, "y": -1, "z": -1, "rgb": -1}
    for name, o, size, typ, count in fields:
        f32 = typ == "F" and size == 4 and count == 1
        if name in ("x", "y", "z") and f32:
            off[name] = o
        if (name == "rgb" and f32) or (name == "rgba" and typ == "U" and size == 4 and count == 1):
            off["rgb"] = o
    return off["x"], off["y"], off["z"], off["rgb"]


def pcd_load_xyzrgb(path):
    """pcl::io::loadPCDFile<pcl::PointXYZRGB>: (float32 [points, 8] rows, header dict)."""
    h = pcd_read(path)
    ox, oy, oz, orgb = field_offsets(h["fields"])
    rows = pc2_to_xyzrgb(h["blob"], h["points"], 1, h["point_step"], h["points"] * h["point_step"], ox, oy, oz, orgb)
    return rows, h


def pcd_write(path, fields, columns, kind, width=None, height=None, comments=True, tabs=False):
    """Write a PCD v0.7 file as pcl::PCDWriter lays it out.  fields: [(name, size, type, count)]; columns: one array
    [n, count] per field (ignored for "_" padding in ascii / compressed bodies)."""
    n = len(columns[0]) if columns else 0
    width = n if width is None else width
    sep = "\t" if tabs else " "
    hdr = []
    if comments:
        hdr.append("# .PCD v0.7 - Point Cloud Data file format")
    hdr.append("VERSION 0.7")
    hdr.append("FIELDS" + sep + sep.join(f[0] for f in fields))
    hdr.append("SIZE" + sep + sep.join(str(f[1]) for f in fields))
    hdr.append("TYPE" + sep + sep.join(f[2] for f in fields))
    hdr.append("COUNT" + sep + sep.join(str(f[3]) for f in fields))
    hdr.append(f"WIDTH {width}")
    if height is not None:
        hdr.append(f"HEIGHT {height}")
    hdr.append("VIEWPOINT 0 0 0 1 0 0 0")
    hdr.append(f"POINTS {n}")
    hdr.append("DATA " + kind)
    head = ("\n".join(hdr) + "\n").encode("ascii")
    cols = [np.ascontiguousarray(np.asarray(c).reshape(n, f[3]).astype(_NP[(f[2], f[1])])) for f, c in zip(fields, columns)]
    if kind == "ascii":
        lines = []
        for i in range(n):
            toks = []
            for f, c in zip(fields, cols):
                for v in c[i]:
                    if f[2] == "F":  # shortest text that reads back to the same float32 / float64
                        toks.append("nan" if np.isnan(v) else np.format_float_scientific(v, unique=True))
                    else:
                        toks.append(str(int(v)))
            lines.append(sep.join(toks))
        body = ("\n".join(lines) + ("\n" if lines else "")).encode("ascii")
    elif kind == "binary":
        body = np.concatenate([c.view(np.uint8).reshape(n, -1) for c in cols], axis=1).tobytes() if n else b""
    else:
        soa = b"".join(c.tobytes() for c in cols)
        comp = lzf_compress(soa)
        body = np.array([len(comp), len(soa)], np.uint32).tobytes() + comp
    open(path, "wb").write(head + body)
