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

The three offline entry points of the reference's ``localization/pose_estimator.py`` are here too, with the same
arguments and result dicts: ``pose_estimator_hloc`` (:138-270), ``pose_refinement`` (:273-377) and
``pose_estimator_iterative`` (:380-612), plus ``get_covisibility_frames`` (:18-42).  They are host control flow around
the two device operators (matcher, pose); ``pose_fn`` lets a caller inject another pose operator (tests pin the
control flow against the reference's own functions that way).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from .. import ops


# COLMAP camera models by parameter layout (colmap/src/colmap/sensor/models.h).  Only models whose image -> camera-plane
# map is implemented below are accepted; anything else raises instead of guessing intrinsics from the wrong slots.
_SINGLE_FOCAL = {'SIMPLE_PINHOLE', 'SIMPLE_RADIAL', 'RADIAL', 'SIMPLE_RADIAL_FISHEYE', 'RADIAL_FISHEYE'}   # (f, cx, cy, ...)
_DOUBLE_FOCAL = {'PINHOLE', 'OPENCV', 'FULL_OPENCV', 'OPENCV_FISHEYE', 'FOV', 'THIN_PRISM_FISHEYE'}        # (fx, fy, cx, cy, ...)
_UNDISTORTABLE = {'SIMPLE_PINHOLE', 'PINHOLE', 'SIMPLE_RADIAL', 'RADIAL', 'OPENCV', 'FULL_OPENCV'}   # lens model implemented below
_MODEL_BY_ID = {0: 'SIMPLE_PINHOLE', 1: 'PINHOLE', 2: 'SIMPLE_RADIAL', 3: 'RADIAL', 4: 'OPENCV', 5: 'OPENCV_FISHEYE',
                6: 'FULL_OPENCV', 7: 'FOV', 8: 'SIMPLE_RADIAL_FISHEYE', 9: 'RADIAL_FISHEYE', 10: 'THIN_PRISM_FISHEYE'}


def _camera_fields(camera):
    if isinstance(camera, dict):
        model, params = camera['model'], camera['params']
    else:
        model = getattr(camera, 'model', None)
        if model is None:
            model = getattr(camera, 'model_name', None)
        params = camera.params
    model = getattr(model, 'name', model)
    if isinstance(model, (int, np.integer)):
        model = _MODEL_BY_ID.get(int(model), str(model))
    return str(model), [float(v) for v in params]


def camera_intrinsics(camera):
    """(fx, fy, cx, cy) of a COLMAP-style camera (dict, namedtuple or pycolmap.Camera-like object with ``model`` /
    ``params``), by the model's parameter layout.  Unknown models raise instead of guessing from the wrong slots."""
    model, params = _camera_fields(camera)
    if model in _SINGLE_FOCAL:
        return params[0], params[0], params[1], params[2]
    if model in _DOUBLE_FOCAL:
        return params[0], params[1], params[2], params[3]
    raise ValueError(f'unknown camera model {model!r} (known: {sorted(_SINGLE_FOCAL | _DOUBLE_FOCAL)})')


def _distortion(model: str, params, x: np.ndarray, y: np.ndarray):
    """(dx, dy) added by the lens model at undistorted camera-plane coordinates (x, y) -- COLMAP's ``Distortion``."""
    r2 = x * x + y * y
    if model == 'SIMPLE_RADIAL':
        rad = params[3] * r2
        return x * rad, y * rad
    if model == 'RADIAL':
        rad = params[3] * r2 + params[4] * r2 * r2
        return x * rad, y * rad
    if model in ('OPENCV', 'FULL_OPENCV'):
        k1, k2, p1, p2 = params[4:8]
        if model == 'OPENCV':
            rad = k1 * r2 + k2 * r2 * r2
        else:
            k3, k4, k5, k6 = params[8:12]
            rad = (1 + k1 * r2 + k2 * r2 * r2 + k3 * r2 ** 3) / (1 + k4 * r2 + k5 * r2 * r2 + k6 * r2 ** 3) - 1
        xy = x * y
        return (x * rad + 2 * p1 * xy + p2 * (r2 + 2 * x * x), y * rad + 2 * p2 * xy + p1 * (r2 + 2 * y * y))
    return np.zeros_like(x), np.zeros_like(y)


def cam_from_img(camera, points2D: np.ndarray) -> np.ndarray:
    """Pixels [n,2] -> undistorted camera-plane coordinates [n,2], float64 (what pycolmap does inside
    ``absolute_pose_estimation`` with the camera it is handed).  The inverse of the lens model is a damped
    fixed-point iteration ``x <- u - d(x)`` run to 1e-12."""
    model, params = _camera_fields(camera)
    fx, fy, cx, cy = camera_intrinsics(camera)
    if model not in _UNDISTORTABLE:
        raise ValueError(f'the lens model of camera model {model!r} is not implemented by the GPU pose operator '
                         f'(implemented: {sorted(_UNDISTORTABLE)}); refusing to ignore its distortion')
    p = np.asarray(points2D, np.float64).reshape(-1, 2)
    u, v = (p[:, 0] - cx) / fx, (p[:, 1] - cy) / fy
    if model in ('SIMPLE_PINHOLE', 'PINHOLE'):
        return np.stack([u, v], 1)
    x, y = u.copy(), v.copy()
    for _ in range(100):
        dx, dy = _distortion(model, params, x, y)
        nx, ny = u - dx, v - dy
        step = max(np.abs(nx - x).max(initial=0.0), np.abs(ny - y).max(initial=0.0))
        x, y = nx, ny
        if step < 1e-12:
            break
    return np.stack([x, y], 1)


_MAX_HYP_PER_LAUNCH = 16384


def absolute_pose_estimation(points2D, points3D, camera, estimation_options: Optional[dict] = None,
                             refinement_options: Optional[dict] = None, device='cuda', seed: int = 0):
    """``ransac`` options honoured: ``max_error`` (pixels; default 12 like pycolmap's absolute-pose default),
    ``min_num_trials`` / ``max_num_trials`` / ``confidence`` (LO-RANSAC's adaptive stopping rule, evaluated between
    fixed-size parallel rounds of hypotheses: after each round the number of trials needed for the requested confidence
    at the best inlier ratio so far is recomputed, as COLMAP does after each improvement), ``min_inlier_ratio``.
    Inputs stay float64: pixels are mapped to the camera plane (with the camera's distortion model) on the host and
    the device estimator runs on float64 correspondences."""
    p2 = np.asarray(points2D, dtype=np.float64).reshape(-1, 2)
    p3 = np.asarray(points3D, dtype=np.float64).reshape(-1, 3)
    n = p2.shape[0]
    if n < 3:
        return None
    r = (estimation_options or {}).get('ransac', {})
    max_error = float(r.get('max_error', 12.0))
    min_trials = int(r.get('min_num_trials', 1000))
    max_trials = max(min_trials, int(r.get('max_num_trials', 100000)))
    confidence = float(r.get('confidence', 0.9999))
    min_inlier_ratio = float(r.get('min_inlier_ratio', 0.01))
    fx, fy, _, _ = camera_intrinsics(camera)
    xy = cam_from_img(camera, p2)
    dev = torch.device(device)
    corr = torch.from_numpy(np.concatenate([xy, p3], 1)[None]).to(dev)
    best, done, rnd = None, 0, 0
    need = min_trials
    while done < need:
        hyp = int(min(_MAX_HYP_PER_LAUNCH, max(need - done, 128)))
        hyp = -(-hyp // 128) * 128                     # whole blocks of the hypothesis kernel
        out = ops.ransac_pnp_corr(corr, 0.5 * (fx + fy), max_error, num_hypotheses=hyp, seed=seed + rnd)
        done += hyp
        rnd += 1
        ni = int(out['num_inliers'][0]) if bool(out['success'][0]) else 0
        if best is None or ni > best[0]:
            best = (ni, out)
        w = best[0] / n
        if w <= 0:
            need = max_trials
        elif w >= 1:
            need = min_trials
        else:
            need = int(min(max_trials, max(min_trials, np.ceil(np.log(1 - confidence) / np.log(1 - w ** 3)))))
    ni, out = best
    if ni < 3 or ni < min_inlier_ratio * n or not bool(out['success'][0]):
        return None
    q = out['qvec'][0].cpu().numpy()
    pose = SimpleNamespace(rotation=SimpleNamespace(quat=q[[1, 2, 3, 0]]), translation=out['tvec'][0].cpu().numpy())
    return {'cam_from_world': pose, 'num_inliers': ni, 'inliers': out['inliers'][0].cpu().numpy()}


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


# ------------------------------------------------------------------------------------------------
# offline entry points (reference localization/pose_estimator.py:18-42, 138-612)
# ------------------------------------------------------------------------------------------------

def get_covisibility_frames(frame_id, all_images, points3D, covisibility_frame: int = 50):
    """Database frames sharing 3-D points with ``frame_id``, most shared points first, at most ``covisibility_frame``
    (reference pose_estimator.py:18-42; ties keep numpy's argsort / argpartition order like the reference)."""
    from collections import defaultdict
    covis = defaultdict(int)
    for pid in all_images[frame_id].point3D_ids:
        if pid == -1:
            continue
        for img_id in points3D[pid].image_ids:
            if img_id != frame_id:
                covis[img_id] += 1
    ids = np.array(list(covis.keys()))
    num = np.array([covis[i] for i in ids])
    if len(ids) <= covisibility_frame:
        return ids[np.argsort(-num)]
    top = np.argpartition(num, -covisibility_frame)[-covisibility_frame:]
    top = top[np.argsort(-num[top])]
    return [ids[i] for i in top]


def _query_from_store(feature_file, qname):
    grp = feature_file[qname]
    rd = lambda k: np.asarray(grp[k][()])
    return {'keypoints': rd('keypoints'), 'scores': rd('scores'), 'descriptors': rd('descriptors').transpose(),
            'image_size': rd('image_size')}


def _camera_from_info(qinfo):
    model, width, height, params = qinfo
    return {'model': model, 'width': width, 'height': height, 'params': params}


def _solve(pose_fn, p2d, p3d, cam, max_error):
    """One call of the pose operator with the reference's post-processing (success flag, wxyz quaternion)."""
    ret = pose_fn(p2d, p3d, cam, estimation_options={'ransac': {'max_error': max_error}}, refinement_options={})
    if ret is None:
        return {'success': False}
    ret['success'] = True
    ret['qvec'] = np.asarray(ret['cam_from_world'].rotation.quat)[[3, 0, 1, 2]]
    ret['tvec'] = ret['cam_from_world'].translation
    return ret


def _log(log_info, text):
    print(text)
    return log_info + text + '\n' if log_info is not None else log_info


def pose_estimator_hloc(qname, qinfo, db_ids, db_images, points3D, feature_file, thresh, image_dir, matcher, log_info=None,
                        query_img_prefix='', db_img_prefix='', pose_fn=None):
    """hloc-style localisation (reference pose_estimator.py:138-270): 2D-3D matches against every retrieved database
    image (3-D points seen in >= 3 images), ONE pose from their union; on failure the pose of the first database image
    is returned as an approximation with ``num_inliers`` 0."""
    import time
    pose_fn = pose_fn or absolute_pose_estimation
    t_start = time.time()
    query = _query_from_store(feature_file, qname)
    cam = _camera_from_info(qinfo)
    best_db_id = db_ids[0]
    best_db_name = db_images[best_db_id].name
    kpts, xyzs, ids3d = [], [], []
    for db_id in db_ids:
        mp3d, mkpq, mp3d_ids, _ = find_2D_3D_matches(query_data=query, db_id=db_id, points3D=points3D, feature_file=feature_file,
                                                     db_images=db_images, matcher=matcher, obs_th=3)
        if mp3d.shape[0] > 0:
            kpts.append(mkpq)
            xyzs.append(mp3d)
            ids3d += list(mp3d_ids)

    def approximate(log_info):
        log_info = _log(log_info, 'Localize {:s} failed, but use the pose of {:s} as approximation'.format(qname, best_db_name))
        return {'qvec': db_images[best_db_id].qvec, 'tvec': db_images[best_db_id].tvec, 'log_info': log_info, 'qname': qname,
                'dbname': best_db_name, 'num_inliers': 0, 'order': -1, 'keypoints_query': np.array([]), 'points3D_ids': [],
                'time': time.time() - t_start}

    if not kpts:
        return approximate(log_info)
    kpts, xyzs = np.vstack(kpts), np.vstack(xyzs)
    ret = _solve(pose_fn, kpts, xyzs, cam, thresh)
    if not ret['success']:
        return approximate(log_info)
    log_info = _log(log_info, 'qname: {:s} localization success with {:d}/{:d} inliers'.format(qname, ret['num_inliers'], xyzs.shape[0]))
    inl = np.asarray(ret['inliers'], bool)
    return {'qvec': ret['qvec'], 'tvec': ret['tvec'], 'log_info': log_info, 'qname': qname, 'dbname': best_db_name,
            'num_inliers': ret['num_inliers'], 'order': -1, 'keypoints_query': np.array([kpts[i] for i in np.nonzero(inl)[0]]),
            'points3D_ids': [ids3d[i] for i in np.nonzero(inl)[0]], 'time': time.time() - t_start}


def pose_refinement(query_data, query_cam, feature_file, db_frame_id, db_images, points3D, matcher, covisibility_frame=50,
                    obs_th=3, opt_th=12, qvec=None, tvec=None, log_info='', pose_fn=None, **kwargs):
    """Covisibility refinement (reference pose_estimator.py:273-377): match the query against the frames covisible with
    ``db_frame_id``, pool the 2D-3D matches (3-D points with >= ``obs_th`` observations), ONE pose with ``opt_th``.
    On failure the incoming ``qvec`` / ``tvec`` are handed back with an all-false inlier list."""
    pose_fn = pose_fn or absolute_pose_estimation
    db_ids = get_covisibility_frames(frame_id=db_frame_id, all_images=db_images, points3D=points3D,
                                     covisibility_frame=covisibility_frame)
    kpq = query_data['keypoints']
    mp3d, mkpq, all_3D_ids = [], [], []
    for db_id in db_ids:
        grp = feature_file[db_images[db_id].name]
        rd = lambda k: np.asarray(grp[k][()])
        points3D_ids = np.asarray(db_images[db_id].point3D_ids)
        if points3D_ids.size == 0:
            print('No 3D points in this db image: ', db_images[db_id].name)
            continue
        matches = np.asarray(feature_matching(query_data=query_data,
                                              db_data={'keypoints': rd('keypoints'), 'scores': rd('scores'),
                                                       'descriptors': rd('descriptors').transpose(), 'image_size': rd('image_size'),
                                                       'db_3D_ids': points3D_ids},
                                              matcher=matcher))
        valid = np.where(matches > -1)[0]
        valid = valid[points3D_ids[matches[valid]] != -1]
        for idx in valid:
            id_3D = points3D_ids[matches[idx]]
            if len(points3D[id_3D].image_ids) < obs_th:
                continue
            mp3d.append(points3D[id_3D].xyz)
            mkpq.append(kpq[idx])
            all_3D_ids.append(id_3D)
    mp3d = np.array(mp3d, float).reshape(-1, 3)
    mkpq = np.array(mkpq, float).reshape(-1, 2) + 0.5
    log_info = _log(log_info, 'Get {:d} covisible frames with {:d} matches from cluster optimization'.format(len(db_ids), mp3d.shape[0]))
    ret = _solve(pose_fn, mkpq, mp3d, query_cam, opt_th)
    ret.update({'mkpq': mkpq, '3D_ids': all_3D_ids, 'db_ids': db_ids, 'log_info': log_info})
    if not ret['success']:
        ret.update({'score_q': [], 'qvec': qvec, 'tvec': tvec, 'inliers': [False for _ in range(mkpq.shape[0])], 'num_inliers': 0,
                    'keypoints_query': np.array([]), 'points3D_ids': []})
        return ret
    inl = np.nonzero(np.asarray(ret['inliers'], bool))[0]
    ret['keypoints_query'] = np.array([mkpq[i] for i in inl])
    ret['points3D_ids'] = [all_3D_ids[i] for i in inl]
    return ret


def pose_estimator_iterative(qname, qinfo, db_ids, db_images, points3D, feature_file, thresh, image_dir, matcher, inlier_th=50,
                             log_info=None, do_covisibility_opt=False, covisibility_frame=50, vis_dir=None, obs_th=0, opt_th=12,
                             gt_qvec=None, gt_tvec=None, query_img_prefix='', db_img_prefix='', pose_fn=None):
    """Iterative localisation over the retrieved database images, best candidate first (reference
    pose_estimator.py:380-612): the first image that yields >= ``inlier_th`` inliers wins (optionally refined over its
    covisible frames); otherwise the best attempt with >= 10 inliers; otherwise the pose of the first database image
    with ``num_inliers`` -1.

    Deviations from the reference, all on paths where the reference itself raises: its ``do_covisibility_opt`` branch
    calls ``pose_refinement`` without the required ``query_data`` argument (:553, :583 -> TypeError) -- the query is
    passed here; its no-candidate fallback indexes ``db_ids[0][0]`` (:601 -> TypeError for integer ids) -- ``db_ids[0]``
    is used; and its ">= 10 inliers" exit reads ``ret`` / ``loc_keypoints_query`` of the LAST loop iteration (:576-594,
    which may be a failed one) -- the recorded best attempt is returned instead."""
    pose_fn = pose_fn or absolute_pose_estimation
    print('qname: ', qname)
    db_name_to_id = {image.name: i for i, image in db_images.items()}
    query = _query_from_store(feature_file, qname)
    cam = _camera_from_info(qinfo)
    best = {'tvec': None, 'qvec': None, 'num_inliers': 0, 'single_num_inliers': 0, 'db_id': -1, 'order': -1, 'qname': qname,
            'optimize': False, 'dbname': db_images[db_ids[0]].name, 'ret_source': '', 'inliers': [],
            'keypoints_query': np.array([]), 'points3D_ids': []}

    def refine(db_frame_id, qvec, tvec):
        return pose_refinement(query_data=query, query_cam=cam, feature_file=feature_file, db_frame_id=db_frame_id,
                               db_images=db_images, points3D=points3D, matcher=matcher, covisibility_frame=covisibility_frame,
                               obs_th=obs_th, opt_th=opt_th, qvec=qvec, tvec=tvec, log_info='', pose_fn=pose_fn)

    for order, db_id in enumerate(db_ids):
        db_name = db_images[db_id].name
        tag = 'qname: {:s} dbname: {:s} ({:d}/{:d})'.format(qname, db_name, order + 1, len(db_ids))
        mp3d, mkpq, mp3d_ids, _ = find_2D_3D_matches(query_data=query, db_id=db_id, points3D=points3D, feature_file=feature_file,
                                                     db_images=db_images, matcher=matcher, obs_th=obs_th)
        if mp3d.shape[0] < 8:
            log_info = _log(log_info, 'qname: {:s} dbname: {:s}({:d}/{:d}) failed because of insufficient 3d points {:d}'.format(
                qname, db_name, order + 1, len(db_ids), mp3d.shape[0]))
            continue
        ret = _solve(pose_fn, mkpq, mp3d, cam, thresh)
        if not ret['success']:
            log_info = _log(log_info, tag + ' failed after matching')
            continue
        inl = np.nonzero(np.asarray(ret['inliers'], bool))[0]
        kq, pids = np.array([mkpq[i] for i in inl]), [mp3d_ids[i] for i in inl]
        if ret['num_inliers'] > best['num_inliers']:
            best.update({'qvec': ret['qvec'], 'tvec': ret['tvec'], 'inlier': ret['inliers'], 'num_inliers': ret['num_inliers'],
                         'dbname': db_name, 'order': order + 1, 'keypoints_query': kq, 'points3D_ids': pids})
        if ret['num_inliers'] < inlier_th:
            log_info = _log(log_info, tag + ' failed insufficient {:d} inliers'.format(ret['num_inliers']))
            continue
        log_info = _log(log_info, tag + ' initialization succeed with {:d} inliers'.format(ret['num_inliers']))
        if do_covisibility_opt:
            ret = refine(db_id, ret['qvec'], ret['tvec'])
            kq, pids = ret['keypoints_query'], ret['points3D_ids']
            if log_info is not None:
                log_info = log_info + ret['log_info']
            log_info = _log(log_info, 'Find {:d} inliers after optimization'.format(ret['num_inliers']))
        best.update({'keypoints_query': kq, 'points3D_ids': pids, 'qvec': ret['qvec'], 'tvec': ret['tvec'],
                     'num_inliers': ret['num_inliers'], 'log_info': log_info})
        return best

    if best['num_inliers'] >= 10:  # 20 for aachen
        if do_covisibility_opt:
            ret = refine(db_name_to_id[best['dbname']], best['qvec'], best['tvec'])
            best.update({'qvec': ret['qvec'], 'tvec': ret['tvec'], 'num_inliers': ret['num_inliers']})
            if ret['success']:
                best.update({'keypoints_query': ret['keypoints_query'], 'points3D_ids': ret['points3D_ids']})
        best['log_info'] = log_info
        return best

    closest = db_images[db_ids[0]]
    log_info = _log(log_info, 'Localize {:s} failed, but use the pose of {:s} as approximation'.format(qname, closest.name))
    best.update({'qvec': closest.qvec, 'tvec': closest.tvec, 'num_inliers': -1, 'log_info': log_info})
    return best
