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
    dev = next((p.device for p in matcher.parameters()), torch.device('cuda'))  # parameter-free matchers (NN) run on cuda
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


def find_2D_3D_matches(query_data: dict, db_id, points3D, feature_file, db_images, matcher, obs_th: int = 0):
    """Same contract as reference ``localization/pose_estimator.py:88-134``: match the query against database image
    ``db_id`` (features read from the h5-like mapping ``feature_file[name][key][()]``, descriptors stored [D, N]),
    keep matches whose database keypoint has a 3-D point observed in at least ``obs_th`` images, and return
    ``(mp3d [n,3] float, mkpq [n,2] float (+0.5 pixel-centre shift), mp3d_ids list, q_ids list)``.
    The matcher call is the device part (``feature_matching``); the filtering is vectorised numpy instead of the
    reference's per-match Python loop (same order: ascending query index)."""
    kpq = query_data['keypoints']
    db_name = db_images[db_id].name
    grp = feature_file[db_name]
    rd = lambda k: np.asarray(grp[k][()])  # h5py datasets and plain numpy arrays both support [()]
    kpdb = rd('keypoints')
    desc_db = rd('descriptors').transpose()
    points3D_ids = np.asarray(db_images[db_id].point3D_ids)
    matches = feature_matching(query_data=query_data,
                               db_data={'keypoints': kpdb, 'scores': rd('scores'), 'descriptors': desc_db,
                                        'db_3D_ids': points3D_ids, 'image_size': rd('image_size')},
                               matcher=matcher)
    matches = np.asarray(matches)
    q_ids = np.nonzero(matches != -1)[0]
    if q_ids.size:
        ids3d = points3D_ids[matches[q_ids]]
        keep = ids3d != -1
        q_ids, ids3d = q_ids[keep], ids3d[keep]
        if obs_th > 0 and q_ids.size:
            keep = np.array([len(points3D[i].image_ids) >= obs_th for i in ids3d], bool)
            q_ids, ids3d = q_ids[keep], ids3d[keep]
    else:
        ids3d = np.zeros((0,), np.int64)
    mp3d = np.array([points3D[i].xyz for i in ids3d], float).reshape(-1, 3)
    mkpq = np.asarray(kpq, float)[q_ids].reshape(-1, 2) + 0.5
    return mp3d, mkpq, [int(i) for i in ids3d], [int(i) for i in q_ids]
