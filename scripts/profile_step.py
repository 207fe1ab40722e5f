"""One registration job on the bench workload (plus the isolated-kernel hooks), short enough to run under ncu.
Usage: python scripts/profile_step.py [points]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from leica_point_cloud_processing_b200 import Engine, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
src, tgt, T_star = synth.make_pair(n, n)
eng = Engine(0)
eng.set_params(max_corr_distance=1.0)
for rep in range(2):
    eng.set_target(tgt)
    eng.set_source(src)
    res = eng.align()
    fit = eng.fitness(res["transform"])
ms_corr, _ = eng.bench_kernel(0, res["transform"], iters=3)
ms_nn, _ = eng.bench_kernel(2, res["transform"], iters=3)
ms_cost, _ = eng.bench_kernel(1, res["transform"], iters=3)
with_fod, _ = synth.add_fod_blobs(synth.apply_rigid(T_star, src).astype(np.float32), n_blobs=20, seed=999)
mask, kept = eng.cloud_difference(with_fod, tgt, 4e-4)
labels, n_clusters = eng.euclidean_clusters(with_fod[mask.astype(bool)], 4e-2, 3, 0)
rgb = np.zeros((n, 8), np.float32)
rgb[:, :3] = src
rgb[:, 3] = 1.0
vox = eng.voxel_grid(rgb, 0.0185)
rows = eng.pointcloud2_to_xyzrgb(rgb, n, 1, 32, 32 * n, 0, 4, 8, 16)
eng.set_params(cost_moments=1)                                                          # the opt-in objective
res_m = eng.align()
eng.set_params(cost_moments=0)
print("outer", res["outer_iterations"], "evals", res["cost_evaluations"], "ms", res["ms_total"], "far", res["corr_far_queries"], "fit", fit,
      "corr_ms", ms_corr, "nn_ms", ms_nn, "cost_ms", ms_cost, "kept", kept, "clusters", n_clusters, "voxels", len(vox),
      "moments-mode ms", res_m["ms_total"], "launches", eng.launch_count())
