"""The index traversal code of the CUDA library (csrc/common.cuh, csrc/knn_search.cuh) compiled for the host and
checked against brute force: exact NN-1 (ungated, gated, seeded, early-exit) and kNN on sheets, lattices with exact
distance ties, duplicates, far-apart clusters and degenerate clouds.  This checks the LOGIC the kernels run (box
search, hierarchical far search, ring termination, pruning margins, tie rule) without a GPU; the GPU parity tests
(-m gpu) check the kernels themselves.  The host build is test infrastructure: the shipped library has no CPU path."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traversal_matches_brute_force(tmp_path):
    exe = str(tmp_path / "emul_search")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off",
                           "-I" + os.path.join(ROOT, "leica_point_cloud_processing_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host_emul", "emul_search.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:]
    assert "0 failures" in out.stdout
    # again with EVERY query answered by the hierarchical far traversal alone (oriented brick slabs, a third of them with
    # an arbitrary direction; thin tilted sheets and thick double sheets among the point sets)
    out = subprocess.run([exe, "far"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:]
    assert "0 failures" in out.stdout
