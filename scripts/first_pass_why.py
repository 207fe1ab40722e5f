"""Why does one strip of the weak-scaling panel cost twice another in the first correspondence pass?  The N = 2 workload with
(a) the panel as it is, (b) the rotation reversed, (c) a smooth panel (no ribs, stringers, rivets), (d) smooth + reversed.
Usage: python scripts/first_pass_why.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from leica_point_cloud_processing_b200 import Engine, synth  # noqa: E402

eng = Engine(0)
eng.set_params(max_corr_distance=1.0)
I = np.eye(4, dtype=np.float32)
N = 2
n = N * 1_000_000
s = N ** 0.5
relief = synth.panel_height
for name, smooth, angle in (("panel, +5 deg", False, 5.0), ("panel, -5 deg", False, -5.0), ("smooth, +5 deg", True, 5.0),
                            ("smooth, -5 deg", True, -5.0)):
    synth.panel_height = (lambda u, s_, length, width: np.zeros_like(u)) if smooth else relief
    src, tgt, _ = synth.make_pair(n, n, length=4.0 * s, width=2.0 * s, angle_deg=angle)
    eng.set_target(tgt)
    q = np.quantile(src[:, 1], np.linspace(0, 1, N + 1))
    for r in range(N):
        part = np.ascontiguousarray(src[(src[:, 1] >= q[r]) & (src[:, 1] <= q[r + 1])])
        eng.set_source(part)
        eng.compute_covariances()
        ms, _ = eng.bench_kernel(3, I, iters=3)
        idx, d2 = eng.nn1(part[::97])
        nn = tgt[idx]
        side = np.mean((part[::97, 2] - nn[:, 2]) > 0)
        print(f"{name} strip {r}: first pass {ms:.3f} ms, offset mean {np.sqrt(d2).mean():.3f} m, queries above their match (z) {100 * side:.0f} %", flush=True)
