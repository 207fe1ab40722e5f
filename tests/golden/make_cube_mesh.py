"""Extract the triangle mesh of the reference's test fixture (test/cube.ply) into a small JSON file.

Run once in the build container (where /root/reference exists); the JSON travels with the repo, the
reference does not.  Reference use of the fixture: test/test_gicp_alignment.cpp:37-46,
test/test_cad_to_pointcloud.cpp:24-55.
"""
import json
import struct
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/test/cube.ply"
raw = open(src, "rb").read()
end = raw.index(b"end_header\n") + len(b"end_header\n")
header = raw[:end].decode().split("\n")
nv = int([l for l in header if l.startswith("element vertex")][0].split()[-1])
nf = int([l for l in header if l.startswith("element face")][0].split()[-1])
off = end
verts = []
for _ in range(nv):
    verts.append(list(struct.unpack_from("<3f", raw, off)))
    off += 12
faces = []
for _ in range(nf):
    (k,) = struct.unpack_from("<B", raw, off)
    off += 1
    faces.append(list(struct.unpack_from("<%di" % k, raw, off)))
    off += 4 * k
assert off == len(raw)
json.dump({"source": "reference test/cube.ply", "vertices": verts, "faces": faces},
          open(__file__.rsplit("/", 1)[0] + "/cube_mesh.json", "w"))
print(nv, nf)
