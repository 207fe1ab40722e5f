"""numpy restatement of the matrix PCL 1.8.1 gicp.hpp computeCovariances hands to its SVD, used by the oracle
hardening tests (CPU) and the GPU covariance parity test: independent of oracle/gicp_oracle.cpp and of the CUDA kernel."""
import numpy as np


def raw_covariances(cloud, knn_idx):
    """cov_i = (1/k) sum_j p p^T - mean mean^T over the k neighbours (products in float32, sums in float64, neighbour
    order), exactly the loop of gicp.hpp.  Returns float64 [n, 3, 3]."""
    p = np.asarray(cloud, np.float32)[:, :3][knn_idx]              # [n, k, 3] float32
    k = knn_idx.shape[1]
    mean = np.zeros((len(p), 3))
    cov = np.zeros((len(p), 3, 3))
    for j in range(k):                                              # neighbour order, as the reference accumulates
        q = p[:, j, :]
        mean += q.astype(np.float64)
        for a in range(3):
            for b in range(a + 1):
                cov[:, a, b] += (q[:, a] * q[:, b]).astype(np.float64)   # float32 product, then widened
    mean /= k
    for a in range(3):
        for b in range(a + 1):
            cov[:, a, b] = cov[:, a, b] / k - mean[:, a] * mean[:, b]
            cov[:, b, a] = cov[:, a, b]
    return cov


def regularised_from_svd(raw, eps=1e-3):
    """U diag(1, 1, eps) U^T with numpy's LAPACK SVD (singular values descending, as Eigen's JacobiSVD orders them)."""
    U, s, _ = np.linalg.svd(raw)
    d = np.array([1.0, 1.0, eps])
    return np.einsum("nik,k,njk->nij", U, d, U), s


def normal_error_bound(s, ulps=64.0):
    """Conditioning-aware bound on |C - C_ref|_max between two correct eigen-solvers of the same symmetric matrix:
    the direction of the smallest singular value moves by ~ |E| / (s2 - s3) under a perturbation E ~ ulps * eps * s1,
    and C = I - (1 - eps) n n^T moves by twice that."""
    gap = np.maximum(s[:, 1] - s[:, 2], 1e-300)
    return 2.0 * ulps * np.finfo(np.float64).eps * s[:, 0] / gap + 1e-13
