"""Device-resident reference frames and the two hot functions of the reference's ``SingleMap3D``.

Scope: ``SingleMap3D.localize_with_ref_frame`` and ``SingleMap3D.match``
(reference ``localization/singlemap3d.py:127-226``) and the ``RefFrame`` accessors they read
(``localization/refframe.py:34-75``).  Same arguments, same result dicts (numpy, same keys) -- but a reference frame's
keypoints / descriptors / scores are uploaded ONCE per (frame, landmark) and stay resident (the reference re-uploads
them on every matcher call, ``singlemap3d.py:143-153``), and the pose comes from the GPU P3P / RANSAC operator instead
of ``pycolmap.absolute_pose_estimation`` (``singlemap3d.py:168-175``).  The id bookkeeping is host numpy, like the reference's.  Map loading (COLMAP
models, ``point3D_desc.npy`` ...) stays outside: the constructor takes already-associated reference frames.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np
import torch

from . import pose_estimator


class RefFrame:
    """A reference frame after ``associate_keypoints_with_point3Ds`` (reference refframe.py:99-129): per keypoint a
    projected position + score, the 3-D point's descriptor / xyz / id / segment id.  Arrays are kept on the host
    (the result dicts are numpy, like the reference's) and mirrored on the device for the matcher."""

    def __init__(self, camera, id: int, keypoints: np.ndarray, descriptors: np.ndarray, xyzs: np.ndarray,
                 point3D_ids: np.ndarray, keypoint_segs: np.ndarray, qvec=None, tvec=None, name: Optional[str] = None,
                 device='cuda'):
        self.camera, self.id, self.qvec, self.tvec, self.name = camera, id, qvec, tvec, name
        self.width, self.height = camera.width, camera.height
        self.image_size = np.array([self.height, self.width])
        self.keypoints = np.asarray(keypoints)            # [n, 3]: u, v, score
        self.descriptors = np.asarray(descriptors)        # [n, D]
        self.xyzs = np.asarray(xyzs)                      # [n, 3]
        self.point3D_ids = np.asarray(point3D_ids)
        self.keypoint_segs = np.asarray(keypoint_segs)
        self.device = torch.device(device)
        self._dev: Dict[object, Dict[str, torch.Tensor]] = {}

    def _pack(self, mask: Optional[np.ndarray]) -> dict:
        sel = slice(None) if mask is None else mask
        return {'point3D_ids': self.point3D_ids[sel], 'keypoints': self.keypoints[sel][:, :2],
                'descriptors': self.descriptors[sel], 'scores': self.keypoints[sel][:, 2], 'xyzs': self.xyzs[sel],
                'camera': self.camera}

    def _device_copy(self, key, data: dict) -> Dict[str, torch.Tensor]:
        if key not in self._dev:
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(self.device)
            self._dev[key] = {'keypoints': t(data['keypoints'])[None], 'descriptors': t(data['descriptors'])[None],
                              'scores': t(data['scores'])[None], 'xyzs': t(data['xyzs'])}
        return self._dev[key]

    def get_keypoints(self) -> dict:
        """reference refframe.py:68-75 (+ '_device': the same arrays as resident tensors)."""
        d = self._pack(None)
        d['_device'] = self._device_copy('all', d)
        return d

    def get_keypoints_by_sid(self, sid: int) -> dict:
        """reference refframe.py:34-42."""
        d = self._pack(self.keypoint_segs == sid)
        d['_device'] = self._device_copy(int(sid), d)
        return d


class SingleMap3D:
    def __init__(self, config: dict, matcher, reference_frames: Dict[int, RefFrame], seg_ref_frame_ids: Dict[int, list],
                 point3D_sids: Dict[int, int], device='cuda', pose_fn: Optional[Callable] = None):
        """``config['localization']['threshold']`` = RANSAC max_error in pixels (configs/*.yaml: 8 or 12);
        ``seg_ref_frame_ids[sid]`` = virtual reference frame ids of landmark ``sid`` (first = best);
        ``point3D_sids[point3D_id]`` = segment id of a 3-D point (``self.point3Ds[v].seg_id`` in the reference);
        ``pose_fn`` defaults to the GPU ``absolute_pose_estimation`` (same signature as pycolmap's)."""
        self.config, self.matcher = config, matcher
        self.reference_frames, self.seg_ref_frame_ids, self.point3D_sids = reference_frames, seg_ref_frame_ids, point3D_sids
        self.device = torch.device(device)
        self.pose_fn = pose_fn or pose_estimator.absolute_pose_estimation

    # -- the matcher call shared by both entry points (reference singlemap3d.py:143-154, 202-213) ----------------
    def _match(self, q_kpts: np.ndarray, q_descs: np.ndarray, q_scores: np.ndarray, q_camera, ref_data: dict) -> np.ndarray:
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))[None].float().to(self.device)
        dev = ref_data.get('_device')
        if dev is None:  # plain dict from a caller: upload like the reference does
            dev = {'keypoints': t(ref_data['keypoints']), 'descriptors': t(ref_data['descriptors']), 'scores': t(ref_data['scores'])}
        ref_cam = ref_data['camera']
        with torch.no_grad():
            out = self.matcher({
                'descriptors0': t(q_descs), 'keypoints0': t(q_kpts), 'scores0': t(q_scores),
                # the reference passes (1, 3, W, H) here -- kept, it changes the keypoint normalisation
                'image_shape0': (1, 3, q_camera.width, q_camera.height),
                'descriptors1': dev['descriptors'], 'keypoints1': dev['keypoints'], 'scores1': dev['scores'],
                'image_shape1': (1, 3, ref_cam.width, ref_cam.height),
            })
        return out['matches0'][0].cpu().numpy()

    def match(self, query_data: dict, ref_data: dict) -> dict:
        """reference singlemap3d.py:196-226."""
        q_kpts = query_data['keypoints']
        indices0 = self._match(q_kpts, query_data['descriptors'], query_data['scores'], query_data['camera'], ref_data)
        valid = indices0 >= 0
        return {'matched_keypoints': q_kpts[valid], 'matched_xyzs': ref_data['xyzs'][indices0[valid]],
                'matched_point3D_ids': ref_data['point3D_ids'][indices0[valid]], 'matched_keypoint_ids': np.where(valid)[0]}

    def localize_with_ref_frame(self, q_frame, q_kpt_ids: np.ndarray, sid: int, semantic_matching: bool = False) -> dict:
        """reference singlemap3d.py:127-194: match the selected query keypoints against the best virtual reference
        frame of landmark ``sid`` and estimate the pose from the 2D-3D matches (+0.5 pixel-centre shift)."""
        ref_frame_id = self.seg_ref_frame_ids[sid][0]
        ref_frame = self.reference_frames[ref_frame_id]
        ref_data = ref_frame.get_keypoints_by_sid(sid=sid) if (semantic_matching and sid > 0) else ref_frame.get_keypoints()
        q_kpt_ids = np.asarray(q_kpt_ids)
        q_descs = q_frame.descriptors[q_kpt_ids]
        q_kpts = q_frame.keypoints[q_kpt_ids, :2]
        q_scores = q_frame.keypoints[q_kpt_ids, 2]
        xyzs, point3D_ids = ref_data['xyzs'], ref_data['point3D_ids']
        ref_sids = np.array([self.point3D_sids[v] for v in point3D_ids])
        indices0 = self._match(q_kpts, q_descs, q_scores, q_frame.camera, ref_data)
        valid = indices0 >= 0
        mkpts, mxyzs = q_kpts[valid], xyzs[indices0[valid]]
        ret = self.pose_fn(mkpts + 0.5, mxyzs, q_frame.camera,
                           estimation_options={'ransac': {'max_error': self.config['localization']['threshold']}},
                           refinement_options={})
        if ret is None:
            ret = {'success': False}
        else:
            ret['success'] = True
            ret['qvec'] = np.asarray(ret['cam_from_world'].rotation.quat)[[3, 0, 1, 2]]
            ret['tvec'] = ret['cam_from_world'].translation
        ret['matched_keypoints'] = mkpts
        ret['matched_keypoint_ids'] = q_kpt_ids[valid]
        ret['matched_xyzs'] = mxyzs
        ret['reference_frame_id'] = ref_frame_id
        ret['matched_point3D_ids'] = point3D_ids[indices0[valid]]
        ret['matched_sids'] = ref_sids[indices0[valid]] if ref_sids.size else ref_sids
        ret['matched_ref_keypoints'] = ref_data['keypoints'][indices0[valid]]
        if not ret['success']:
            ret['num_inliers'] = 0
            ret['inliers'] = np.zeros(shape=(mkpts.shape[0],), dtype=bool)
        return ret
