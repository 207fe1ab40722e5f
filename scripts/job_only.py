"""One registration job (set_clouds + align + fitness) at a given size, twice; for launch lists under ncu.
Usage: python scripts/job_only.py [points]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leica_point_cloud_processing_b200 import Engine, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
src, tgt, T_star = synth.make_pair(n, n)
eng = Engine(0)
eng.set_params(max_corr_distance=1.0)
for rep in range(2):
    t0 = time.perf_counter()
    eng.set_target(tgt)
    t1 = time.perf_counter()
    eng.set_source(src)
    t2 = time.perf_counter()
    eng.compute_covariances()
    t3 = time.perf_counter()
    res = eng.align()
    t4 = time.perf_counter()
    fit = eng.fitness(res["transform"])
    t5 = time.perf_counter()
    print("n", n, "index_t %.3f index_s %.3f cov %.3f align %.3f fitness %.3f ms" % tuple(1e3 * (b - a) for a, b in
          ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5))), "outer", res["outer_iterations"], "evals", res["cost_evaluations"],
          "corr_ms", res["ms_corr"], "launches", eng.launch_count())
