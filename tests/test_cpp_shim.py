"""The C++ drop-in (include/GICPAlignment_b200.hpp) against the reference's own gtests, restated in
tests/cpp/test_shim.cpp.  CPU: the header compiles as C++14 (the reference's standard, CMakeLists.txt:4-6) and links
against the C-ABI library.  GPU: the binary runs the reference's three GICPAlignment tests plus the removeFromCloud
case on the cube fixture, and the transform it reports is checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "leica_point_cloud_processing_b200")


def build(tmp_path, std="c++14"):
    exe = str(tmp_path / "test_shim")
    subprocess.check_call(["g++", f"-std={std}", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"), "-o", exe, "-L" + PKG, "-lgicp_b200",
                           "-Wl,-rpath," + PKG])
    return exe


def test_shim_compiles_and_links(tmp_path):
    exe = build(tmp_path)
    out = subprocess.run(["nm", "-u", exe], capture_output=True, text=True).stdout
    for sym in ("gicpb_create", "gicpb_align", "gicpb_set_clouds", "gicpb_fitness",
                "gicpb_transform_cloud", "gicpb_cloud_difference", "gicpb_euclidean_clusters", "gicpb_voxel_grid",
                "gicpb_pointcloud2_to_xyzrgb", "gicpb_pcd_load_xyzrgb", "gicpb_normals", "gicpb_cloud_resolution"):
        assert sym in out
    # the reference's public surface is all there (include/GICPAlignment.h:47-145)
    hdr = open(os.path.join(ROOT, "include", "GICPAlignment_b200.hpp")).read()
    for name in ("GICPAlignment(CloudPtr target_cloud, CloudPtr source_cloud, bool use_covariances)", "void run()",
                 "void iterate()", "void undo()", "Matrix4f getFineTransform()", "void getAlignedCloud(CloudPtr",
                 "void applyTFtoCloud(CloudPtr", "void setSourceCloud(CloudPtr", "void setTargetCloud(CloudPtr",
                 "void setMaxIterations(int", "void setTfEpsilon(double", "void setMaxCorrespondenceDistance(int",
                 "void setRANSACOutlierTh(int", "bool transform_exists_;"):
        assert name in hdr, name


@pytest.mark.gpu
def test_reference_gtests_through_the_cpp_shim(tmp_path, cube_pair, oracle):
    from leica_point_cloud_processing_b200 import synth
    from oracle.oracle import default_params
    src, tgt, T = cube_pair
    src.astype(np.float32).tofile(str(tmp_path / "source.f32"))
    tgt.astype(np.float32).tofile(str(tmp_path / "target.f32"))
    from oracle import cloud_io as oio
    s32 = src.astype(np.float32)
    oio.pcd_write(str(tmp_path / "source.pcd"), [("x", 4, "F", 1), ("y", 4, "F", 1), ("z", 4, "F", 1)],
                  [s32[:, 0], s32[:, 1], s32[:, 2]], "binary_compressed")
    # the pipeline-order check: a 60 k panel scan at its offset pose with FOD blobs on it, and the CAD cloud
    p_src, p_tgt, T_star = synth.make_pair(60_000, 60_000)
    fod_aligned, _ = synth.add_fod_blobs(synth.apply_rigid(T_star, p_src), n_blobs=6)
    scan = synth.apply_rigid(np.linalg.inv(T_star), fod_aligned).astype(np.float32)
    scan.tofile(str(tmp_path / "scan.f32"))
    p_tgt.astype(np.float32).tofile(str(tmp_path / "cad.f32"))
    exe = build(tmp_path)
    run = subprocess.run([exe, str(tmp_path / "source.f32"), str(tmp_path / "target.f32"), str(tmp_path / "source.pcd"),
                          str(tmp_path / "scan.f32"), str(tmp_path / "cad.f32")],
                         capture_output=True, text=True, timeout=600)
    print(run.stdout[-4000:], run.stderr[-2000:])
    assert run.returncode == 0
    res = {}
    for line in run.stdout.splitlines():
        if line.startswith("RESULT "):
            parts = line.split()
            res[parts[1]] = parts[2:]
    assert res["failed"] == ["0"]
    assert res["pcd_points"] == [str(len(src))]  # loadPCDFile (src/load_and_publish_clouds.cpp:75) through the shim
    assert res["fod_clusters"][0] == "3"      # test/test_fod_detector.cpp:70 ASSERT_EQ(num_of_fods, 3)
    # testRun parameters (gate 5, tf_eps 5e-4): same transform as the oracle, inside the north_star tolerances
    T_gpu = np.array([float(v) for v in res["run_transform"]]).reshape(4, 4)
    ref = oracle.align(src, tgt, default_params(max_corr_distance=5.0, transformation_epsilon=5e-4))
    diag = float(np.linalg.norm(tgt.max(0) - tgt.min(0)))
    assert synth.rotation_error_rad(T_gpu, ref["T"]) <= 1e-4
    assert synth.translation_error(T_gpu, ref["T"]) <= 1e-5 * diag
    assert synth.rotation_error_rad(T_gpu, T) <= 2e-3  # and it is the yaw the fixture applied
    fit_ref = oracle.fitness(src, tgt, ref["T"])
    assert abs(float(res["run_fitness"][0]) - fit_ref) <= 1e-4 * abs(fit_ref) + 1e-12
    # ---- one LeicaStateMachine run through the shims, stage by stage against the oracle doing the same stages ----------
    def rows32(xyz):
        r = np.zeros((len(xyz), 8), np.float32)
        r[:, :3] = xyz
        r[:, 3] = 1.0
        return r
    t_res, s_res = oracle.resolution(p_tgt), oracle.resolution(scan)
    assert np.allclose([float(v) for v in res["pipeline_resolution"]], [t_res, s_res], rtol=1e-12)
    leaf = 10.0 * max(t_res, s_res)
    o_src = oracle.voxel_grid(rows32(scan), leaf)[:, :3].copy()
    o_tgt = oracle.voxel_grid(rows32(p_tgt), leaf)[:, :3].copy()
    assert [int(v) for v in res["pipeline_downsampled"]] == [len(o_src), len(o_tgt)]
    o_al = oracle.align(o_src, o_tgt, default_params(max_corr_distance=1.0))
    T_pipe = np.array([float(v) for v in res["pipeline_transform"]]).reshape(4, 4)
    p_diag = float(np.linalg.norm(o_tgt.max(0) - o_tgt.min(0)))
    assert synth.rotation_error_rad(T_pipe, o_al["T"]) <= 1e-4
    assert synth.translation_error(T_pipe, o_al["T"]) <= 1e-5 * p_diag
    moved = oracle.transform(o_al["T"], o_src)
    th = 4e-3 * 3
    o_mask, o_kept = oracle.difference(moved, o_tgt, th)
    assert int(res["pipeline_difference"][0]) == o_kept
    o_labels, o_nc = oracle.euclidean_clusters(moved[o_mask.astype(bool)], th * 100, 3, 0)
    assert int(res["pipeline_fods"][0]) == o_nc
    assert [int(v) for v in res["pipeline_fods"][1:]] == [int((o_labels == k).sum()) for k in range(o_nc)]
    o_nrm, o_fin = oracle.normals(o_tgt, 4.0 * leaf)
    assert int(res["pipeline_normals"][0]) == o_fin
    ok = np.isfinite(o_nrm[:, 0])
    o_sum = float((o_nrm[ok, 0].astype(np.float64) + 2.0 * o_nrm[ok, 1] + 3.0 * o_nrm[ok, 2] + o_nrm[ok, 3]).sum())
    assert abs(float(res["pipeline_normals"][1]) - o_sum) <= 1e-4 * max(1.0, abs(o_sum))
    # default parameters (gate 0.04): same as the oracle too
    T_def = np.array([float(v) for v in res["applytf_transform"]]).reshape(4, 4)
    ref_def = oracle.align(src, tgt, default_params())
    assert synth.rotation_error_rad(T_def, ref_def["T"]) <= 1e-4
    assert synth.translation_error(T_def, ref_def["T"]) <= 1e-5 * diag
