"""Loop-closure re-fusion of submaps (ElevationMapping::updateGlobalMap, ElevationMapping.cpp:773-905):
the host-side outer loops -- pose update, kd-tree neighbour selection -- around the two device calls
gem_transform_cloud and gem_refuse_submaps.  `backend` is anything with transform_cloud / refuse_submaps methods: a
gem_b200.ElevationMap (device tensors) or the oracle adapter of the tests (numpy arrays)."""
from __future__ import annotations

import numpy as np


def neighbours(centres, i: int, radius: float = 25.0):
    """KdTreeFLANN::radiusSearch on the submap centres (ElevationMapping.cpp:821-838): indices within `radius` of centre i,
    nearest first (squared float distances; ties by index: DEFINITION, FLANN's tie order is unspecified)"""
    c = np.asarray(centres, np.float32)
    d2 = ((c[:, 0] - c[i, 0]) ** 2 + (c[:, 1] - c[i, 1]) ** 2).astype(np.float32)
    idx = np.flatnonzero(d2 <= np.float32(radius) * np.float32(radius))
    return idx[np.lexsort((idx, d2[idx]))].tolist()


def update_global_map(backend, submaps, old_poses, new_poses, centres, resolution: float, radius: float = 25.0,
                      compat: bool = True):
    """submaps: list of (n, 8) PointXYZRGBICT arrays (backend's array type), modified and possibly shortened; poses: 4x4
    arrays (trajectory_ and optGlobalMapLoc_).  Returns (submaps, fused cell count)."""
    K = len(submaps)
    for i in range(1, K):                                   # :796-812 (submap 0 keeps its pose)
        T = (np.asarray(new_poses[i], np.float32) @ np.linalg.inv(np.asarray(old_poses[i], np.float32))).astype(np.float32)
        backend.transform_cloud(submaps[i], T)
    total = 0
    for i in range(K):                                      # :815-891
        nb = neighbours(centres, i, radius)
        if len(nb) > 2:                                     # :841
            for j in nb[1:]:                                # :843: skip the nearest (the submap itself)
                if j == i:
                    continue   # DEFINITION: a tie in distance can put i later in the list; fusing a map with itself is skipped
                nn, no, fused = backend.refuse_submaps(submaps[j], submaps[i], resolution, compat)
                submaps[j], submaps[i] = submaps[j][:nn], submaps[i][:no]
                total += fused
    return submaps, total
