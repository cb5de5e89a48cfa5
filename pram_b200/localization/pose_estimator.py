"""Pose operator -- a GPU replacement for the call the reference makes into pycolmap:

    ret = pycolmap.absolute_pose_estimation(points2D, points3D, camera,
                                            estimation_options={'ransac': {'max_error': th, ...}},
                                            refinement_options={})

(reference localization/singlemap3d.py:168-175, :324-333, :454; tracker.py:211; pose_estimator.py:213,
338, 452).  Same argument meaning and return convention: ``None`` on failure, else a dict with
``cam_from_world`` (``.rotation.quat`` in **xyzw** order, ``.translation``), ``num_inliers`` and ``inliers``
(bool[n]); the reference reorders the quaternion to wxyz itself (singlemap3d.py:180).
``feature_matching`` mirrors reference pose_estimator.py:45-86 (packing of query / database features for
the matcher and the id remap).

Parity with pycolmap 0.6.1 is unpinned (the wheel is not available); the estimator is
P3P + RANSAC with local optimisation and a Cauchy-weighted refinement on the device (csrc/ransac.cu).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from .. import ops


def camera_intrinsics(camera):
    """(fx, fy, cx, cy) from a COLMAP-style camera (dict, namedtuple or pycolmap.Camera-like object with
    ``model`` / ``params``); distortion parameters are ignored."""
    model = camera['model'] if isinstance(camera, dict) else getattr(camera, 'model', getattr(camera, 'model_name', None))
    params = camera['params'] if isinstance(camera, dict) else camera.params
    params = [float(v) for v in params]
    model = getattr(model, 'name', model)
    if str(model) in ('PINHOLE', 'OPENCV', 'FULL_OPENCV', 'OPENCV_FISHEYE'):
        return params[0], params[1], params[2], params[3]
    return params[0], params[0], params[1], params[2]


def absolute_pose_estimation(points2D, points3D, camera, estimation_options: Optional[dict] = None,
                             refinement_options: Optional[dict] = None, device='cuda', seed: int = 0):
    p2 = np.asarray(points2D, dtype=np.float64).reshape(-1, 2)
    p3 = np.asarray(points3D, dtype=np.float64).reshape(-1, 3)
    n = p2.shape[0]
    if n < 3:
        return None
    r = (estimation_options or {}).get('ransac', {})
    max_error = float(r.get('max_error', 12.0))
    # fixed-size parallel sampling: at least min_num_trials, at most 16384 hypotheses per call
    trials = int(min(16384, max(1024, r.get('min_num_trials', 1000))))
    fx, fy, cx, cy = camera_intrinsics(camera)
    dev = torch.device(device)
    k = torch.from_numpy(p2).to(dev).float()[None]
    x = torch.from_numpy(p3).to(dev).float()[None]
    m = torch.arange(n, device=dev)[None]
    out = ops.ransac_pnp(k, m, x, fx, fy, cx, cy, max_error, pixel_shift=0.0, num_hypotheses=trials, seed=seed)
    if not bool(out['success'][0]):
        return None
    q = out['qvec'][0].cpu().numpy()
    pose = SimpleNamespace(rotation=SimpleNamespace(quat=q[[1, 2, 3, 0]]), translation=out['tvec'][0].cpu().numpy())
    return {'cam_from_world': pose, 'num_inliers': int(out['num_inliers'][0]), 'inliers': out['inliers'][0].cpu().numpy()}


def feature_matching(query_data: dict, db_data: dict, matcher) -> np.ndarray:
    """Reference pose_estimator.py:45-86: match query features against one database image, optionally
    restricted to keypoints with a 3-D point, and remap the matches to database keypoint ids."""
    dev = next(matcher.parameters()).device
    db_3D_ids = db_data.get('db_3D_ids')
    if db_3D_ids is None:
        valid_ids = None
        sel = slice(None)
    else:
        valid_ids = np.nonzero(np.asarray(db_3D_ids) != -1)[0]
        if valid_ids.size == 0:
            return np.zeros((query_data['keypoints'].shape[0],), dtype=int) - 1
        sel = valid_ids

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a))[None].float().to(dev)

    data = {
        'keypoints0': t(query_data['keypoints']), 'scores0': t(query_data['scores']),
        'descriptors0': t(query_data['descriptors']),
        'image0': torch.empty((1, 1) + tuple(query_data['image_size'])[::-1], device='meta'),
        'keypoints1': t(db_data['keypoints'][sel]), 'scores1': t(db_data['scores'][sel]),
        'descriptors1': t(db_data['descriptors'][sel]),
        'image1': torch.empty((1, 1) + tuple(db_data['image_size'])[::-1], device='meta'),
    }
    with torch.no_grad():
        matches = matcher(data)['matches0'][0].cpu().numpy()
    if valid_ids is not None:
        ok = matches >= 0
        matches[ok] = valid_ids[matches[ok]]
    return matches
