"""How many queries take the hierarchical far path, per kind of search, on the bench workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from leica_point_cloud_processing_b200 import Engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
src, tgt, T_star = synth.make_pair(n, n)
eng = Engine(0)
eng.set_params(max_corr_distance=1.0)
eng.set_target(tgt); eng.set_source(src)
eng.compute_covariances(); print("knn(src) far", eng.last_far_queries())
res = eng.align(); print("align far total", res["corr_far_queries"], "outer", res["outer_iterations"])
T = res["transform"]
eng.bench_kernel(0, T, iters=1); print("corr @converged unseeded far", eng.last_far_queries())
eng.bench_kernel(0, np.eye(4, dtype=np.float32), iters=1); print("corr @identity far", eng.last_far_queries())
eng.fitness(T); print("fitness far", eng.last_far_queries())
idx, d2 = eng.nn1(synth.apply_rigid(T_star, src)); print("nn1 far", eng.last_far_queries(), "d2 mean", float(d2.mean()), "d2 p99", float(np.quantile(d2, 0.99)), "max", float(d2.max()))
print("grid", eng.grid_info(0))
