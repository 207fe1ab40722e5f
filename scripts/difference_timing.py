"""Where the time of a 10 M-point cloud difference goes: host (pageable / pinned) against device-resident clouds."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from leica_point_cloud_processing_b200 import Engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
src, tgt, T_star = synth.make_pair(n, n, length=12.0, width=4.0)
aligned = synth.apply_rigid(T_star, src)
with_fod, _ = synth.add_fod_blobs(aligned, n_blobs=20, length=12.0, width=4.0)
eng = Engine(0)
def t(label, fn, reps=3):
    for r in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
        print(f"{label} run {r}: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
    return out
t("pageable numpy", lambda: eng.cloud_difference(with_fod, tgt, 4e-4))
pf, pt = torch.from_numpy(with_fod).pin_memory(), torch.from_numpy(tgt).pin_memory()
t("pinned host   ", lambda: eng.cloud_difference(pf, pt, 4e-4))
df, dt_ = pf.cuda(), pt.cuda()
t("device        ", lambda: eng.cloud_difference(df, dt_, 4e-4))
t0 = time.perf_counter(); eng.difference_set_subtract(dt_) if hasattr(eng, "difference_set_subtract") else None; print("set_subtract", (time.perf_counter() - t0) * 1e3)
# whole rows both ways: pcl::transformPointCloud of 10 M PointXYZRGB rows held in pageable / pinned host memory / on the device
rows = np.zeros((n, 8), np.float32); rows[:, :3] = src; rows[:, 3] = 1
Tm = np.eye(4, dtype=np.float32); Tm[0, 3] = 0.5
t("transform 320 MB pageable", lambda: eng.transform_cloud(Tm, rows, out=rows))
pr = torch.from_numpy(rows).pin_memory()
t("transform 320 MB pinned  ", lambda: eng.transform_cloud(Tm, pr, out=pr))
dr = pr.cuda()
t("transform 320 MB device  ", lambda: eng.transform_cloud(Tm, dr, out=dr))
