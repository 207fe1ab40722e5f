"""Cost-kernel time with and without the persisting-L2 window, fp64 / fp32 Mahalanobis, at several sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from leica_point_cloud_processing_b200 import Engine, synth
for n in (1_000_000, 4_000_000):
    src, tgt, T_star = synth.make_pair(n, n)
    for fp32 in (0, 1):
        for persist in (1,):
            eng = Engine(0)
            eng.set_params(max_corr_distance=1.0, mahalanobis_fp32=fp32, l2_persist=persist)
            eng.set_target(tgt); eng.set_source(src)
            res = eng.align()
            ms, _ = eng.bench_kernel(1, res["transform"], iters=50)
            print(f"n {n} fp32 {fp32} persist {persist}: cost eval {ms*1e3:.1f} us, align {res['ms_total']:.2f} ms (cost total {res['ms_cost']:.2f}) evals {res['cost_evaluations']}")
            eng.close()
