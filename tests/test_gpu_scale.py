"""BASELINE.json configurations at their FULL sizes on one B200, checked through properties that do not need the
oracle to repeat the whole job:
  config 2  1 M / 1 M panel: whole alignment against the oracle (north_star tolerances)
  config 3  10 M / 10 M fuselage section: recovers the known offset T*; exact-NN spot check against the oracle's kd-tree
            on a query sample; fitness consistent with the 1 mm range noise
  config 4  FOD difference on the aligned 10 M pair with injected blobs: mask bit-exact against the oracle at full
            size for the launch threshold, FOD points kept, surface points dropped
Sizes can be cut with GICPB_TEST_SCALE (fraction of the point counts) for a quick run; GICPB_TEST_FULL_ORACLE=1 adds
the whole-alignment comparison with the oracle at 10 M points (minutes of host time)."""
import os
import time

import numpy as np
import pytest

from leica_point_cloud_processing_b200 import synth

pytestmark = pytest.mark.gpu
SCALE = float(os.environ.get("GICPB_TEST_SCALE", "1.0"))
ROT_TOL, TRANS_REL_TOL, FIT_REL_TOL = 1e-4, 1e-5, 1e-4


@pytest.fixture(scope="module")
def fast_oracle():
    from oracle.oracle import Oracle
    return Oracle(fast=True)


def test_config2_1M_matches_oracle(engine, fast_oracle):
    from oracle.oracle import default_params
    n = int(1_000_000 * SCALE)
    src, tgt, T_star = synth.make_pair(n, n)
    engine.set_params(max_corr_distance=1.0, max_iterations=100, transformation_epsilon=4e-3, cell_size=0.0)
    engine.set_target(tgt)
    engine.set_source(src)
    res = engine.align()
    fit = engine.fitness(res["transform"])
    ref = fast_oracle.align(src, tgt, default_params(max_corr_distance=1.0))
    fit_ref = fast_oracle.fitness(src, tgt, ref["T"])
    diag = float(np.linalg.norm(tgt.max(0) - tgt.min(0)))
    assert res["converged"] == 1
    assert synth.rotation_error_rad(res["transform"], ref["T"]) <= ROT_TOL
    assert synth.translation_error(res["transform"], ref["T"]) <= TRANS_REL_TOL * diag
    assert abs(fit - fit_ref) <= FIT_REL_TOL * fit_ref
    assert res["outer_iterations"] == ref["outer_iterations"]
    # and it is the offset the generator applied, as far as the reference's stopping rule goes: the outer loop ends
    # when no entry of T moved by more than rotation_epsilon 2e-3 / transformation_epsilon 4e-3 (reference
    # src/GICPAlignment.cpp:29), so the last step's size bounds what is left
    assert synth.rotation_error_rad(res["transform"], T_star) <= 6e-3
    assert synth.translation_error(res["transform"], T_star) <= 2e-2


@pytest.fixture(scope="module")
def pair_10M():
    n = int(10_000_000 * SCALE)
    src, tgt, T_star = synth.make_pair(n, n, length=12.0, width=4.0)
    return src, tgt, T_star


def test_config3_10M_alignment_properties(engine, fast_oracle, pair_10M):
    src, tgt, T_star = pair_10M
    engine.set_params(max_corr_distance=1.0, max_iterations=100, transformation_epsilon=4e-3, cell_size=0.0)
    t0 = time.perf_counter()
    engine.set_target(tgt)
    engine.set_source(src)
    res = engine.align()
    fit = engine.fitness(res["transform"])
    dt = time.perf_counter() - t0
    print(f"10M/10M on one GPU: {dt * 1e3:.1f} ms end to end from host clouds, outer {res['outer_iterations']}, "
          f"evals {res['cost_evaluations']}, fitness {fit:.3e}")
    assert res["converged"] == 1 and res["status"] == 0
    rot, tr = synth.rotation_error_rad(res["transform"], T_star), synth.translation_error(res["transform"], T_star)
    pre = engine.fitness(np.eye(4, dtype=np.float32))
    fit_true = engine.fitness(T_star.astype(np.float32))
    print(f"vs T*: rot {rot:.2e} rad, trans {tr:.2e} m; fitness before {pre:.3e} after {fit:.3e} at T* {fit_true:.3e}")
    # The panel is nearly a surface of revolution (sliding along it is held only by 15-20 mm ribs and rivets) and the
    # reference stops once no entry of T moves by more than 2e-3 / 4e-3, so "equals T*" is not a sharp property of
    # the algorithm here; "is much better than the start and in the neighbourhood of T*" is.
    assert rot <= 1.5e-2 and tr <= 5e-2
    assert 1e-7 < fit < 1e-4 and pre > 50 * fit and fit < 5 * fit_true
    # the sharp statement at full size: the same pose, iteration count and fitness as the oracle.  The oracle needs
    # ~4 minutes of all host threads for 10 M points, so this part runs only with GICPB_TEST_FULL_ORACLE=1 (a logged
    # run is committed as profiles/r01_scale_parity_10M.log); config 2 above does the same at 1 M on every run.
    if os.environ.get("GICPB_TEST_FULL_ORACLE", "0") == "1":
        _compare_with_oracle(fast_oracle, src, tgt, res, fit)
    _nn_spot_check(engine, fast_oracle, src, tgt, T_star)


def _compare_with_oracle(fast_oracle, src, tgt, res, fit):
    from oracle.oracle import default_params
    t0 = time.perf_counter()
    ref = fast_oracle.align(src, tgt, default_params(max_corr_distance=1.0))
    fit_ref = fast_oracle.fitness(src, tgt, ref["T"])
    print(f"oracle on {fast_oracle.num_threads()} host threads: {time.perf_counter() - t0:.1f} s, outer {ref['outer_iterations']}")
    diag = float(np.linalg.norm(tgt.max(0) - tgt.min(0)))
    assert synth.rotation_error_rad(res["transform"], ref["T"]) <= ROT_TOL
    assert synth.translation_error(res["transform"], ref["T"]) <= TRANS_REL_TOL * diag
    assert abs(fit - fit_ref) <= FIT_REL_TOL * fit_ref
    assert res["outer_iterations"] == ref["outer_iterations"]


def _nn_spot_check(engine, fast_oracle, src, tgt, T_star):
    # exact NN at full size: a query sample against the oracle's kd-tree over all 10 M target points, bit for bit
    rng = np.random.default_rng(7)
    q = synth.apply_rigid(T_star, src[rng.choice(len(src), 20_000, replace=False)])
    q[:2000] += rng.normal(0, 0.05, (2000, 3)).astype(np.float32)   # some queries centimetres off the surface
    q[2000:2100] += np.float32(3.0)                                  # and some metres away / outside the grid
    idx, d2 = engine.nn1(q)
    oi, od = fast_oracle.nn1(tgt, q)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)


def test_config4_fod_difference_10M_bit_exact(engine, fast_oracle, pair_10M):
    src, tgt, T_star = pair_10M
    aligned = synth.apply_rigid(T_star, src)
    with_fod, is_fod = synth.add_fod_blobs(aligned, n_blobs=20, length=12.0, width=4.0)
    thr = 4e-3 * 0.1  # launch value of voxelize_factor (reference launch file :16, src/LeicaStateMachine.cpp:188)
    t0 = time.perf_counter()
    mask, kept = engine.cloud_difference(with_fod, tgt, thr)
    dt = time.perf_counter() - t0
    print(f"difference {len(with_fod)} vs {len(tgt)} points: {dt * 1e3:.1f} ms from host clouds, kept {kept}")
    assert kept == int(mask.sum())
    # every blob point sits >= 10 mm - 15 mm ... above the surface: those farther than sqrt(thr) = 2 cm must be kept;
    # surface points (1 mm noise, 2 mm spacing) must all go
    assert mask[~is_fod].sum() == 0
    assert mask[is_fod].sum() > 0
    om, ok = fast_oracle.difference(with_fod, tgt, thr)
    assert ok == kept
    assert np.array_equal(mask, om)
    # idempotence: the kept points, run again, are all kept; the dropped ones are all dropped
    mask2, kept2 = engine.cloud_difference(with_fod[mask.astype(bool)], tgt, thr)
    assert kept2 == kept and mask2.all()
    # SURVEY 8f row 1 on the same data (src/LeicaStateMachine.cpp:200-205): the kept points cluster into the injected
    # blobs, label for label as the oracle's restatement of pcl::EuclideanClusterExtraction says
    diff_cloud = with_fod[mask.astype(bool)]
    labels, n_clusters = engine.euclidean_clusters(diff_cloud, thr * 100, 3, 0)
    olab, onc = fast_oracle.euclidean_clusters(diff_cloud, thr * 100, 3, 0)
    assert n_clusters == onc and 1 <= n_clusters <= 20
    assert np.array_equal(labels, olab)


def test_voxel_grid_10M_bit_exact(engine, fast_oracle, pair_10M):
    """SURVEY 8f row 3 at the size of config 3: Filter::downsampleCloud of the 10 M-point scan as PointXYZRGB rows with
    the leaf the pipeline uses (10 x cloud resolution, src/LeicaStateMachine.cpp:61-65); centroids bit for bit, and
    the size-independent properties: every point lands in exactly one voxel, centroids lie inside their voxel."""
    src, _, _ = pair_10M
    rng = np.random.default_rng(3)
    cloud = np.zeros((len(src), 8), np.float32)
    cloud[:, :3] = src
    cloud[:, 3] = 1.0
    cloud[:, 4] = rng.integers(0, 2 ** 32, len(src), dtype=np.uint64).astype(np.uint32).view(np.float32)
    engine.set_target(src)
    leaf = np.float32(10.0 * engine.cloud_resolution(0))
    t0 = time.perf_counter()
    out = engine.voxel_grid(cloud, leaf)
    dt = time.perf_counter() - t0
    print(f"voxel grid {len(cloud)} -> {len(out)} points, leaf {float(leaf) * 1e3:.1f} mm: {dt * 1e3:.1f} ms from host clouds")
    ref = fast_oracle.voxel_grid(cloud, leaf)
    assert out.shape == ref.shape
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    inv = np.float32(1.0) / leaf
    vox = np.floor(out[:, :3] * inv).astype(np.int64)
    assert len(np.unique(vox, axis=0)) == len(out)            # one centroid per voxel, each inside its own voxel
    assert np.all(np.diff(((vox - vox.min(0)) * [1, (vox[:, 0].max() - vox[:, 0].min() + 1),
                                                  (vox[:, 0].max() - vox[:, 0].min() + 1) * (vox[:, 1].max() - vox[:, 1].min() + 1)]).sum(1)) > 0)
