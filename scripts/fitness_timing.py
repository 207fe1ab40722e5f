"""Wall time of gicpb_fitness (seeded by the last correspondence pass) on the bench workload.
Usage: python scripts/fitness_timing.py [points]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from leica_point_cloud_processing_b200 import Engine, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
src, tgt, _ = synth.make_pair(n, n)
eng = Engine(0)
eng.set_params(max_corr_distance=1.0)
eng.set_clouds(tgt, src)
res = eng.align()
T = res["transform"]
for _ in range(5):
    eng.fitness(T)
ts = []
for _ in range(50):
    t0 = time.perf_counter()
    f = eng.fitness(T)
    ts.append(time.perf_counter() - t0)
print(f"n {n} fitness {f:.6e}: median {1e3 * np.median(ts):.3f} ms, min {1e3 * min(ts):.3f} ms")
