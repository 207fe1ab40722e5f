"""Sweep the density target of the grid (points per cell) on the bench workload: phase times per setting."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from leica_point_cloud_processing_b200 import Engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
src, tgt, T_star = synth.make_pair(n, n)
d_src, d_tgt = torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda()
for ppc in [float(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "1.5,2,3,4,6".split(","))]:
    eng = Engine(0)
    eng.set_params(max_corr_distance=1.0, points_per_cell=ppc)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); eng.set_target(d_tgt); eng.set_source(d_src); t_idx = time.perf_counter() - t0
        t0 = time.perf_counter(); eng.compute_covariances(); t_cov = time.perf_counter() - t0
        t0 = time.perf_counter(); res = eng.align(); t_al = time.perf_counter() - t0
        t0 = time.perf_counter(); fit = eng.fitness(res["transform"]); t_fit = time.perf_counter() - t0
    ms_corr, _ = eng.bench_kernel(0, res["transform"], iters=5)
    ms_cold, _ = eng.bench_kernel(0, np.eye(4, dtype=np.float32), iters=2)
    gi = eng.grid_info(0)
    print(f"ppc {ppc}: h {gi['cell_size']*1e3:.2f} mm cells {gi['n_cells_occupied']} bricks {gi['n_bricks_occupied']} | index {t_idx*1e3:.2f} cov {t_cov*1e3:.2f} "
          f"align {t_al*1e3:.2f} (corr {res['ms_corr']:.2f}, far {res['corr_far_queries']}) fitness {t_fit*1e3:.2f} | corr@conv {ms_corr:.3f} corr@identity {ms_cold:.3f} ms")
    eng.close()
