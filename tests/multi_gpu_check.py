"""Run under torchrun (one rank per GPU): the sharded engine against the oracle.  Source points are sharded by rank,
target covariances are computed in chunks and all-gathered, every cost evaluation all-reduces 14 doubles; the result
must be the single-GPU / oracle result on every rank.  Used by tests/test_multi_gpu.py (needs >= 2 GPUs)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from leica_point_cloud_processing_b200 import Engine, synth
    from leica_point_cloud_processing_b200.distributed import env_rank_world, init_engine_comm
    from oracle.oracle import Oracle, default_params

    rank, world, local_rank = env_rank_world()
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    orc = Oracle()
    eng = Engine(local_rank)
    fused = init_engine_comm(eng, rank, world)
    print(f"rank {rank}/{world}: cost sums reduced over peer memory inside the cost kernel: {fused}", flush=True)

    mesh = json.load(open(os.path.join(ROOT, "tests", "golden", "cube_mesh.json")))
    src = orc.sample_mesh(np.array(mesh["vertices"], np.float32), np.array(mesh["faces"], np.int32), 5000)
    tgt = orc.transform(orc.rotation_rpy(0.0, 0.0, 0.175), src)
    cases = [("cube gate 5", src, tgt, dict(max_corr_distance=5.0, transformation_epsilon=5e-4)),
             ("cube default gate", src, tgt, dict())]
    ps, pt, _ = synth.make_pair(60_000, 60_000)
    cases.append(("panel 60k gate 1", ps, pt, dict(max_corr_distance=1.0)))
    failures = 0
    # first with the fused peer-memory reduction (when the box allows it), then the same cases on ncclAllReduce
    modes = ["peer", "nccl"] if fused else ["nccl"]
    for mode in modes:
        if mode == "nccl" and fused:
            eng.peer_disable()
        for name, s, t, prm in cases:
            name = f"[{mode}] {name}"
            eng.set_params(**{**dict(max_corr_distance=4e-2, transformation_epsilon=4e-3), **prm})
            if mode == "peer":
                eng.set_clouds(t, s)   # target covariance chunk on the second stream beside the source index, then all-gather
            else:
                eng.set_target(t)
                eng.set_source(s)
            res = eng.align()
            fit = eng.fitness(res["transform"])
            ref = orc.align(s, t, default_params(**prm))
            fit_ref = orc.fitness(s, t, ref["T"])
            diag = float(np.linalg.norm(t.max(0) - t.min(0)))
            rot = synth.rotation_error_rad(res["transform"], ref["T"])
            tr = synth.translation_error(res["transform"], ref["T"])
            ok = (res["converged"] == 1 and rot <= 1e-4 and tr <= 1e-5 * diag and abs(fit - fit_ref) <= 1e-4 * abs(fit_ref)
                  and res["outer_iterations"] == ref["outer_iterations"])
            # every rank must hold the same answer
            Tt = torch.from_numpy(res["transform"].copy()).cuda()
            T0 = Tt.clone()
            dist.broadcast(T0, 0)
            same = bool(torch.equal(Tt, T0))
            print(f"rank {rank}/{world} {name}: rot {rot:.2e} trans {tr:.2e} fit {fit:.6e} (oracle {fit_ref:.6e}) outer "
                  f"{res['outer_iterations']}/{ref['outer_iterations']} same_on_all_ranks {same} -> "
                  f"{'ok' if ok and same else 'FAILED'}", flush=True)
            failures += 0 if (ok and same) else 1
    eng.close()
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
