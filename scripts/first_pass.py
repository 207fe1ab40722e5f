"""Times the correspondence pass and NN-1 alone at the initial pose (queries 2-20 cm off the surface: far search) and
at the converged pose (near search) on the bench workload.  Usage: python scripts/first_pass.py [points]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from leica_point_cloud_processing_b200 import Engine, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
src, tgt, T_star = synth.make_pair(n, n)
eng = Engine(0)
eng.set_params(max_corr_distance=1.0, cell_size=float(os.environ.get("CELL", "0")))
eng.set_target(tgt)
eng.set_source(src)
res = eng.align()
I = np.eye(4, dtype=np.float32)
for name, T, which in (("initial pose", I, 3), ("converged pose", res["transform"], 0)):
    ms_corr, _ = eng.bench_kernel(which, T, iters=5)
    far = eng.last_far_queries()
    ms_nn, _ = eng.bench_kernel(2, T, iters=5)
    print(f"{name}: correspondence pass {ms_corr:.3f} ms, NN-1 only {ms_nn:.3f} ms, far queries {far}")
for w, nm in ((0, "target"), (1, "source")):
    gi = eng.grid_info(w)
    print(nm, "grid: h", gi["cell_size"], "dims", gi["dims"], "points/cell", gi["n_indexed"] / max(gi["n_cells_occupied"], 1))
print("align ms", res["ms_total"], "corr ms", res["ms_corr"], "outer", res["outer_iterations"])
