"""Frame-to-frame tracking step: ``Tracker.track_last_frame`` of the reference
(``localization/tracker.py:162-233``) -- match the current frame against the previous, already localised frame
(whose keypoints carry 3-D points), keep matches with a valid 3-D point, estimate the pose from the 2D-3D pairs.
Same arguments and result keys as the reference; the matcher and the pose operator are the device ones
(``localization.matchers.*``, ``pose_estimator.absolute_pose_estimation`` instead of pycolmap).
Tracking is sequential per sequence, so multi-GPU sharding is per sequence in this mode."""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch

from . import pose_estimator


class Tracker:
    def __init__(self, config: dict, matcher, device='cuda', pose_fn: Optional[Callable] = None):
        self.config, self.matcher = config, matcher
        self.device = torch.device(device)
        self.pose_fn = pose_fn or pose_estimator.absolute_pose_estimation

    def track_last_frame(self, curr_frame, last_frame) -> dict:
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))[None].float().to(self.device)
        curr_kpts, last_kpts = curr_frame.keypoints[:, :2], last_frame.keypoints[:, :2]
        with torch.no_grad():
            indices = self.matcher({
                'descriptors0': t(curr_frame.descriptors), 'keypoints0': t(curr_kpts), 'scores0': t(curr_frame.keypoints[:, 2]),
                'image_shape0': (1, 3, curr_frame.camera.width, curr_frame.camera.height),
                'descriptors1': t(last_frame.descriptors), 'keypoints1': t(last_kpts), 'scores1': t(last_frame.keypoints[:, 2]),
                'image_shape1': (1, 3, last_frame.camera.width, last_frame.camera.height),
            })['matches0'][0].cpu().numpy()
        q_ids = np.nonzero(indices >= 0)[0]
        ref = indices[q_ids]
        pids = np.asarray(last_frame.point3D_ids)[ref]
        keep = pids >= 0                      # keypoints of the last frame without a 3-D point carry id -1
        q_ids, ref, pids = q_ids[keep], ref[keep], pids[keep]
        matched_kpts, matched_xyzs = curr_kpts[q_ids], np.asarray(last_frame.xyzs)[ref]
        ret = self.pose_fn(matched_kpts + 0.5, matched_xyzs, curr_frame.camera,
                           estimation_options={'ransac': {'max_error': self.config['localization']['threshold']}},
                           refinement_options={})
        if ret is None:
            ret = {'success': False}
        else:
            ret['success'] = True
            ret['qvec'] = np.asarray(ret['cam_from_world'].rotation.quat)[[3, 0, 1, 2]]
            ret['tvec'] = ret['cam_from_world'].translation
        ret['matched_keypoints'] = matched_kpts
        ret['matched_keypoint_ids'] = q_ids
        ret['matched_ref_keypoints'] = last_kpts[ref]
        ret['matched_xyzs'] = matched_xyzs
        ret['matched_point3D_ids'] = pids
        ret['matched_sids'] = np.asarray(last_frame.seg_ids)[ref]
        ret['reference_frame_id'] = last_frame.reference_frame_id
        ret['matched_scene_name'] = last_frame.matched_scene_name
        return ret
