"""FODDetector (reference include/FODDetector.h, src/FODDetector.cpp) and Filter::downsampleCloud (reference
src/Filter.cpp:91-105) on the CUDA engine: the two callers either side of the registration path.

Same method names and argument meaning as the reference classes, so the parity tests read like
test/test_fod_detector.cpp and test/test_filter.cpp.  Clouds are float32 arrays [n, 3 | 4 | 8]; with 8 columns a row
has the memory layout of pcl::PointXYZRGB (x, y, z, 1, packed rgba bits, 3 pad words).
"""
import logging

import numpy as np

from ._capi import Engine

log = logging.getLogger("leica_point_cloud_processing_b200")


class FODDetector:
    """reference src/FODDetector.cpp:21-26: FODDetector(cloud, cluster_tolerance, min_fod_points)."""

    def __init__(self, cloud, cluster_tolerance, min_fod_points, device=0, engine=None):
        self._engine = engine if engine is not None else Engine(device)
        self.cloud_ = cloud
        self.cluster_indices_ = []
        self.setClusterTolerance(cluster_tolerance)
        self.setMinFODpoints(min_fod_points)

    def setClusterTolerance(self, tolerance):  # src/FODDetector.cpp:28-38: 0 is invalid -> default 4e-3
        if tolerance == 0:
            log.warning("FODDetector: invalid tolerance value: %f", tolerance)
            self.cluster_tolerance_ = 4e-3
        else:
            self.cluster_tolerance_ = tolerance
        log.info("FODDetector: cluster tolerance set to: %f", self.cluster_tolerance_)

    def setMinFODpoints(self, min_fod_points):  # :40-43; setMinClusterSize(int) truncates the double
        self.min_cluster_size_ = min_fod_points

    def clusterPossibleFODs(self):  # :45-58 -> pcl::EuclideanClusterExtraction::extract
        n = int(self.cloud_.shape[0])
        labels, nc = self._engine.euclidean_clusters(self.cloud_, self.cluster_tolerance_,
                                                     min_size=int(self.min_cluster_size_), max_size=0)
        order = np.argsort(labels, kind="stable")  # ascending point index inside each cluster, as PointIndices holds them
        counts = np.bincount(labels[labels >= 0], minlength=nc) if n else np.zeros(0, np.int64)
        start = int((labels < 0).sum())
        self.cluster_indices_ = []
        for k in range(nc):
            self.cluster_indices_.append(order[start:start + int(counts[k])].astype(np.int32))
            start += int(counts[k])
        log.info("cluster_indices_size: %d", len(self.cluster_indices_))

    def getFODIndices(self):  # :112-115
        return list(self.cluster_indices_)

    def fodIndicesToPointCloud(self, fod_cloud_array):  # :60-78: appends one cloud per cluster, returns their number
        n_fods = 0
        for idx in self.cluster_indices_:
            sel = idx
            if hasattr(self.cloud_, "is_cuda"):
                import torch
                sel = torch.as_tensor(idx.astype(np.int64), device=self.cloud_.device)
            fod_cloud_array.append(self.cloud_[sel])
            log.info("cluster_size: %d", len(idx))
            n_fods += 1
        return n_fods


def downsample_cloud(cloud, leaf_size, engine=None, device=0):
    """Filter::downsampleCloud (reference src/Filter.cpp:91-105): pcl::VoxelGrid with leaf (leaf_size,)*3, every field
    downsampled; the centroids come in ascending voxel index, PCL's output order."""
    eng = engine if engine is not None else Engine(device)
    log.info("Downsample cloud with leaf_size : %f", leaf_size)
    return eng.voxel_grid(cloud, leaf_size)
