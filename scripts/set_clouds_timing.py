"""Wall time of indexing both clouds + covariances: separate calls against gicpb_set_clouds (device-resident clouds)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from leica_point_cloud_processing_b200 import Engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
src, tgt, _ = synth.make_pair(n, n)
d_src, d_tgt = torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda()
eng = Engine(0)
eng.set_params(max_corr_distance=1.0)
def sep():
    eng.set_target(d_tgt); eng.set_source(d_src); eng.compute_covariances()
def both():
    eng.set_clouds(d_tgt, d_src)
for name, fn in (("separate", sep), ("set_clouds", both), ("separate", sep), ("set_clouds", both)):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    print(f"{os.environ.get('GICPB_AUX_PRIO', '-')}/{os.environ.get('GICPB_CARVEOUT', '-')} n {n} {name}: median {np.median(ts) * 1e3:.3f} ms  min {min(ts) * 1e3:.3f} ms", flush=True)
