"""First correspondence pass (identity pose) against the size of the initial offset and the size of the panel:
is the far search bound by how far the queries are off the surface, or by the working set?
Usage: python scripts/first_pass_offsets.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from leica_point_cloud_processing_b200 import Engine, synth  # noqa: E402

eng = Engine(0)
eng.set_params(max_corr_distance=1.0)
I = np.eye(4, dtype=np.float32)
for n, scale, angle in ((1_000_000, 1.0, 5.0), (1_000_000, 1.0, 10.0), (1_000_000, 1.0, 15.0), (8_000_000, 8 ** 0.5, 5.0),
                        (8_000_000, 8 ** 0.5, 1.77)):
    src, tgt, _ = synth.make_pair(n, n, length=4.0 * scale, width=2.0 * scale, angle_deg=angle)
    eng.set_target(tgt)
    eng.set_source(src)
    eng.compute_covariances()
    ms, _ = eng.bench_kernel(3, I, iters=3)
    far = eng.last_far_queries()
    idx, d2 = eng.nn1(src[::97])
    d = np.sqrt(d2)
    print(f"n {n} panel {4 * scale:.1f} x {2 * scale:.1f} m, {angle} deg: first pass {ms:.3f} ms = {1e6 * ms / n:.2f} ns per query, far {far}, "
          f"offset mean {d.mean():.3f} m p95 {np.quantile(d, 0.95):.3f} m", flush=True)
