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
                "gicpb_pointcloud2_to_xyzrgb", "gicpb_pcd_load_xyzrgb"):
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
    exe = build(tmp_path)
    run = subprocess.run([exe, str(tmp_path / "source.f32"), str(tmp_path / "target.f32"), str(tmp_path / "source.pcd")],
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
    # default parameters (gate 0.04): same as the oracle too
    T_def = np.array([float(v) for v in res["applytf_transform"]]).reshape(4, 4)
    ref_def = oracle.align(src, tgt, default_params())
    assert synth.rotation_error_rad(T_def, ref_def["T"]) <= 1e-4
    assert synth.translation_error(T_def, ref_def["T"]) <= 1e-5 * diag
