"""Phase times of one registration job at several sizes (second run of each size: buffers allocated, clouds in HBM)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from leica_point_cloud_processing_b200 import Engine, synth
sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["2000000", "5000000", "10000000"])]
eng = Engine(0)
eng.set_params(max_corr_distance=1.0)
for n in sizes:
    s = (n / 10_000_000) ** 0.5
    L, W = (4.0, 2.0) if n <= 2_000_000 else (12.0 * s, 4.0 * s)
    src, tgt, T_star = synth.make_pair(n, n, length=L, width=W)
    d_src, d_tgt = torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda()
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); eng.set_target(d_tgt); t_tgt = time.perf_counter() - t0
        t0 = time.perf_counter(); eng.set_source(d_src); t_src = time.perf_counter() - t0
        t0 = time.perf_counter(); eng.compute_covariances(); t_cov = time.perf_counter() - t0
        t0 = time.perf_counter(); res = eng.align(); t_al = time.perf_counter() - t0
        t0 = time.perf_counter(); fit = eng.fitness(res["transform"]); t_fit = time.perf_counter() - t0
    ms_cost, _ = eng.bench_kernel(1, res["transform"], iters=10)
    ms_corr, _ = eng.bench_kernel(0, res["transform"], iters=3)
    gi = eng.grid_info(0)
    tot = t_tgt + t_src + t_cov + t_al + t_fit
    print(f"n {n}: total {tot*1e3:.1f} ms | index {t_tgt*1e3:.2f}+{t_src*1e3:.2f} cov {t_cov*1e3:.2f} align {t_al*1e3:.2f} "
          f"(corr {res['ms_corr']:.2f}, outer {res['outer_iterations']}, evals {res['cost_evaluations']}, far {res['corr_far_queries']}) "
          f"fitness {t_fit*1e3:.2f} | cost eval {ms_cost*1e3:.1f} us, corr@conv {ms_corr:.3f} ms | h {gi['cell_size']*1e3:.2f} mm", flush=True)
