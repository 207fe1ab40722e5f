"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: agreeing on the NCCL unique id over torch.distributed,
contiguous source sharding, and the fact that all-reducing the 14 raw partial sums of each shard and then applying the
host scaling reproduces the single-process objective and gradient (what csrc/engine.cu run_cost + cost_from_sums do
with ncclAllReduce on the GPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rotation_gradient(x, Rs):
    cphi, sphi = np.cos(x[3]), np.sin(x[3])
    cth, sth = np.cos(x[4]), np.sin(x[4])
    cpsi, spsi = np.cos(x[5]), np.sin(x[5])
    dphi = np.array([[0, sphi * spsi + cphi * cpsi * sth, cphi * spsi - cpsi * sphi * sth],
                     [0, -cpsi * sphi + cphi * spsi * sth, -cphi * cpsi - sphi * spsi * sth],
                     [0, cphi * cth, -cth * sphi]])
    dth = np.array([[-cpsi * sth, cpsi * cth * sphi, cphi * cpsi * cth],
                    [-spsi * sth, cth * sphi * spsi, cphi * cth * spsi],
                    [-cth, -sphi * sth, -cphi * sth]])
    dpsi = np.array([[-cth * spsi, -cphi * cpsi - sphi * spsi * sth, cpsi * sphi - cphi * spsi * sth],
                     [cpsi * cth, -cphi * spsi + cpsi * sphi * sth, sphi * spsi + cphi * cpsi * sth],
                     [0, 0, 0]])
    return [float(np.trace(d @ Rs)) for d in (dphi, dth, dpsi)]


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from leica_point_cloud_processing_b200.distributed import exchange_unique_id, shard_range
    from oracle.oracle import Oracle

    uid = exchange_unique_id(lambda: bytes((7 * i + 3) % 256 for i in range(128)), rank)
    assert uid == bytes((7 * i + 3) % 256 for i in range(128))

    data = np.load(os.path.join(tmpdir, "case.npz"))
    src, tgt, idx, maha, x = data["src"], data["tgt"], data["idx"], data["maha"], data["x"]
    orc = Oracle()
    lo, hi = shard_range(len(src), rank, world)
    T = orc.apply_state(x)
    pp = orc.transform(T, src[lo:hi])
    j = idx[lo:hi]
    ok = j >= 0
    res = (pp[ok] - tgt[j[ok]]).astype(np.float64)          # float32 subtraction, then widened (as PCL)
    t = np.einsum("nij,nj->ni", maha[lo:hi][ok], res)
    sums = np.zeros(14)
    sums[0] = (res * t).sum()
    sums[1:4] = t.sum(0)
    sums[4:13] = np.einsum("ni,nj->ij", src[lo:hi][ok].astype(np.float64), t).reshape(-1)
    sums[13] = ok.sum()
    buf = torch.from_numpy(sums)
    dist.all_reduce(buf)                                      # the 14-double all-reduce
    s = buf.numpy()
    m = s[13]
    f = s[0] / m
    g = np.zeros(6)
    g[:3] = s[1:4] * 2.0 / m
    g[3:] = _rotation_gradient(x, s[4:13].reshape(3, 3) * 2.0 / m)
    np.save(os.path.join(tmpdir, f"out{rank}.npy"), np.concatenate([[f], g, [m]]))
    dist.destroy_process_group()


def test_sharded_sums_reproduce_the_objective(tmp_path, oracle, cube_pair):
    src, tgt, _ = cube_pair
    cov_s, cov_t = oracle.covariances(src), oracle.covariances(tgt)
    cnt, idx, d2, maha = oracle.correspondences(src, tgt, cov_s, cov_t, np.eye(4), 0.2)
    x = np.array([0.01, -0.02, 0.015, 0.02, 0.01, 0.12])
    np.savez(tmp_path / "case.npz", src=src, tgt=tgt, idx=idx, maha=maha, x=x)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    valid = np.nonzero(idx >= 0)[0].astype(np.int32)
    f, g = oracle.cost(src, tgt, valid, idx[valid], maha, x)
    for rank in range(2):
        out = np.load(tmp_path / f"out{rank}.npy")
        assert out[7] == cnt
        assert abs(out[0] - f) <= 1e-12 * abs(f)
        assert np.allclose(out[1:7], g, rtol=1e-10, atol=1e-13)


class _FakeEngine:
    """Stands in for Engine in the handle exchange of distributed.enable_peer_reduction (no GPU on this box)."""

    def __init__(self, rank, export_fails=False, import_fails=False):
        self.rank, self.export_fails, self.import_fails = rank, export_fails, import_fails
        self.imported, self.disabled = None, False

    def peer_export(self):
        if self.export_fails:
            raise RuntimeError("no peer access")
        return bytes([self.rank]) * 64

    def peer_import(self, handles):
        if self.import_fails:
            raise RuntimeError("cudaIpcOpenMemHandle failed")
        self.imported = list(handles)

    def peer_disable(self):
        self.disabled = True


def _peer_worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from leica_point_cloud_processing_b200.distributed import enable_peer_reduction

    # every rank can map every other: all enable, and each holds the handles in RANK order
    ok_all = _FakeEngine(rank)
    assert enable_peer_reduction(ok_all, rank, world) is True
    assert ok_all.imported == [bytes([r]) * 64 for r in range(world)] and not ok_all.disabled
    # one rank cannot export, another cannot import: nobody may use the fused path (a half-enabled job would hang)
    for kw in (dict(export_fails=(rank == 1)), dict(import_fails=(rank == 0))):
        eng = _FakeEngine(rank, **kw)
        assert enable_peer_reduction(eng, rank, world) is False
        assert eng.disabled
    # what the last block of the cost kernel does with the slots (cost.cu): sums in rank order on every rank give
    # bit-identical results everywhere, whichever rank's partials arrived first
    rng = np.random.default_rng(100 + rank)
    mine = rng.standard_normal(14) * 10.0 ** rng.integers(-8, 8, 14)
    slots = [torch.zeros(14, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(slots, torch.from_numpy(mine))
    total = np.zeros(14)
    for r in range(world):
        total += slots[r].numpy()
    np.save(os.path.join(tmpdir, f"peer{rank}.npy"), total)
    dist.destroy_process_group()


def test_peer_reduction_is_all_or_nothing_and_rank_ordered(tmp_path):
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_peer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "peer0.npy"), np.load(tmp_path / "peer1.npy")
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
