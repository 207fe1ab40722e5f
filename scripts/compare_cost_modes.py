"""Compares the two objective-evaluation modes of the engine on the same correspondences (one B200):
cost_moments=0 (one cost-kernel pass per evaluation, float T*p as PCL) vs cost_moments=1 (74 second-order moments
per outer iteration, evaluations on the host), and the alignments they lead to."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leica_point_cloud_processing_b200 import Engine, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
src, tgt, T_star = synth.make_pair(n, n)
eng = Engine(0)
out = {}
for mode in (0, 1):
    eng.set_params(max_corr_distance=1.0, cost_moments=mode)
    eng.set_target(tgt); eng.set_source(src)
    eng.correspondences(np.eye(4, dtype=np.float32))
    vals = []
    for x in ([0, 0, 0, 0, 0, 0], [0.01, -0.005, 0.004, 0.02, -0.03, 0.05], [1e-4, 0, 0, 0, 1e-4, 0]):
        f, g = eng.cost(np.array(x, dtype=np.float64))
        vals.append((f, g))
    res = eng.align()
    out[mode] = (vals, res)
for (f0, g0), (f1, g1) in zip(out[0][0], out[1][0]):
    print(f"f kernel {f0:.15e} moments {f1:.15e} rel {abs(f0-f1)/abs(f0):.2e} | g max abs diff {np.abs(g0-g1).max():.2e} (|g| {np.abs(g0).max():.2e})")
r0, r1 = out[0][1], out[1][1]
print("outer", r0["outer_iterations"], r1["outer_iterations"], "evals", r0["cost_evaluations"], r1["cost_evaluations"],
      "ms", r0["ms_total"], r1["ms_total"])
print("rot diff", synth.rotation_error_rad(r0["transform"], r1["transform"]), "trans diff",
      synth.translation_error(r0["transform"], r1["transform"]))
for r in (r0, r1):
    print("vs T*: rot", synth.rotation_error_rad(r["transform"], T_star), "trans", synth.translation_error(r["transform"], T_star),
          "fitness", eng.fitness(r["transform"]))
