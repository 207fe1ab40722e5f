"""First correspondence pass of the weak-scaling workloads (N M / N M points, iso-density panels), strip by strip on ONE GPU:
the source is cut into N strips of equal counts along y (what the ranks of an N-GPU job own) and each strip is searched
alone against the whole target.  Prints ns per query against the strip's mean offset: the cost model behind the
cost-balanced shards (GridIndex::build, ShardCostProbe).  Usage: python scripts/first_pass_strips.py [N ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from leica_point_cloud_processing_b200 import Engine, synth  # noqa: E402

eng = Engine(0)
eng.set_params(max_corr_distance=1.0)
I = np.eye(4, dtype=np.float32)
for N in [int(a) for a in sys.argv[1:]] or [2, 4, 8]:
    n = N * 1_000_000
    s = N ** 0.5
    src, tgt, _ = synth.make_pair(n, n, length=4.0 * s, width=2.0 * s)
    eng.set_target(tgt)
    q = np.quantile(src[:, 1], np.linspace(0, 1, N + 1))
    tot = 0.0
    for r in range(N):
        part = np.ascontiguousarray(src[(src[:, 1] >= q[r]) & (src[:, 1] <= q[r + 1])])
        eng.set_source(part)
        eng.compute_covariances()
        ms, _ = eng.bench_kernel(3, I, iters=3)
        idx, d2 = eng.nn1(part[::97])
        d = np.sqrt(d2)
        tot += ms
        print(f"N {N} strip {r}: {len(part)} queries, first pass {ms:.3f} ms = {1e6 * ms / len(part):.2f} ns per query, offset mean {d.mean():.3f} m "
              f"p90 {np.quantile(d, 0.9):.3f} m", flush=True)
    print(f"N {N}: sum {tot:.3f} ms, mean {tot / N:.3f} ms", flush=True)
