"""A small pass through every kernel of the library with cross-checks between equivalent call paths (set_clouds against
the separate calls, aligned against byte-wise PointCloud2 gathers).  Written for compute-sanitizer
(`compute-sanitizer --tool memcheck python scripts/all_kernels_small.py`); that tool is closed on this GPU pool, so it
serves as a quick functional pass instead."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from leica_point_cloud_processing_b200 import Engine, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30_000
src, tgt, T_star = synth.make_pair(n, n + 777)
eng = Engine(0)
eng.set_params(max_corr_distance=1.0)
eng.set_clouds(tgt, src)
res = eng.align()
fit = eng.fitness(res["transform"])
eng.set_target(tgt); eng.set_source(src)
res2 = eng.align()
assert np.array_equal(res["transform"], res2["transform"])
idx, d2 = eng.nn1(src, res["transform"], 0.5)
ki, kd = eng.knn(1)
with_fod, _ = synth.add_fod_blobs(synth.apply_rigid(T_star, src).astype(np.float32), n_blobs=5, seed=999)
mask, kept = eng.cloud_difference(with_fod, tgt, 4e-4)
labels, n_clusters = eng.euclidean_clusters(with_fod[mask.astype(bool)], 4e-2, 3, 0)
rgb = np.zeros((n, 8), np.float32); rgb[:, :3] = src; rgb[:, 3] = 1.0
vox = eng.voxel_grid(rgb, 0.05)
rows = eng.pointcloud2_to_xyzrgb(rgb, n, 1, 32, 32 * n, 0, 4, 8, 16)
odd = np.zeros((n, 13), np.uint8); odd[:, :12] = src.view(np.uint8).reshape(n, 12)
rows2 = eng.pointcloud2_to_xyzrgb(odd, n, 1, 13, 13 * n, 0, 4, 8, -1)
assert np.array_equal(rows[:, :3], rows2[:, :3])
res_r = eng.cloud_resolution(0); val, nv = eng.normal_validity(0, 4 * res_r)
eng.set_params(cost_moments=1); res_m = eng.align(); eng.set_params(cost_moments=0, mahalanobis_fp32=1); res_f = eng.align()
moved = eng.transform_cloud(res["transform"], src)
print("ok: outer", res["outer_iterations"], "evals", res["cost_evaluations"], "fit", fit, "kept", kept, "clusters", n_clusters,
      "voxels", len(vox), "valid normals", nv, "launches", eng.launch_count())
eng.close()
