"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the disagreement-map rasterisation and the exploration reward.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this.

Follows (numpy executes the assignments, so numpy's own fancy-indexing semantics are the oracle):
  * grid size / grid coordinates     experimenting_env/utils/projection_utils.py:681-699
  * `_safe_assign`                   experimenting_env/utils/projection_utils.py:614-631
  * score gather, centre + inflate   experimenting_env/utils/projection_utils.py:760-785
  * reward = map.sum() / 1000        experimenting_env/envs/env_habitat3.py:75-84

PARITY UNPINNED by reference tests (the reference has none for this function); the assignments below are
the reference's statements verbatim in structure, executed by the same numpy.
"""
from __future__ import annotations

import numpy as np
import torch


def grid_geometry(lower_bound, upper_bound, map_scale=0.025):
    """-> (H, W, off_x, off_y) exactly as projection_utils.py:683-699 computes them."""
    H = int(abs(upper_bound[2] - lower_bound[2]) / map_scale)
    W = int(abs(upper_bound[0] - lower_bound[0]) / map_scale)
    off_x = int(np.floor(lower_bound[0] / map_scale))
    off_y = int(np.ceil(lower_bound[2] / map_scale))
    return H, W, off_x, off_y


def _safe_assign(im_map, y_idx, x_idx, value):
    try:
        im_map[y_idx, x_idx] = value
    except IndexError:
        valid_idx1 = np.logical_and(x_idx >= 0, x_idx < im_map.shape[1])
        valid_idx2 = np.logical_and(y_idx >= 0, y_idx < im_map.shape[0])
        valid_idx = np.logical_and(valid_idx1, valid_idx2)
        if valid_idx.sum() == 0:
            return
        try:
            if isinstance(value, np.ndarray):
                value = value[valid_idx]
            im_map[y_idx[valid_idx], x_idx[valid_idx]] = value
        except (IndexError, ValueError):
            return


def disagreement_map(points: np.ndarray, object_ids, cosine_distances: dict, lower_bound, upper_bound,
                     map_scale=0.025):
    """points float32 [P,3]; object_ids [P]; cosine_distances {object id: 0-d torch tensor} -> (map f32 [H,W], reward)."""
    H, W, off_x, off_y = grid_geometry(lower_bound, upper_bound, map_scale)
    disagreement_mat = np.zeros((H, W), np.float32)
    if len(points):
        grid_x = (points[:, 0] / map_scale).astype(int) - off_x
        grid_y = (points[:, 2] / map_scale).astype(int) - off_y
        cosine_distance_per_point = [cosine_distances[int(k)] for k in object_ids]
        if len(cosine_distance_per_point):
            cosine_distance_per_point = torch.stack(cosine_distance_per_point)
            valid = cosine_distance_per_point > 0
            _safe_assign(disagreement_mat, grid_y[valid], grid_x[valid], cosine_distance_per_point[valid])
            for i in range(-1, 2):
                for j in range(-1, 2):
                    if i != j:
                        _safe_assign(disagreement_mat, grid_y[valid] + i, grid_x[valid] + j,
                                     cosine_distance_per_point[valid])
    reward = disagreement_mat.sum() / 1000.0
    return disagreement_mat, float(reward)
