"""Frame-to-frame tracking: the reference's ``Tracker`` (``localization/tracker.py``) --
``run`` (:37-127), ``verify_and_update`` / ``update_current_frame`` (:129-160), ``track_last_frame`` (:162-233),
``track_last_frame_fast`` (:235-311) and ``match_frame`` (:313-338).  Match the current frame against the previous, already
localised frame (whose keypoints carry 3-D points), keep matches with a valid 3-D point, estimate the pose from the
2D-3D pairs; refine through the map when fewer than 256 inliers survive.  Same constructor order, arguments and result
keys as the reference; the matcher and the pose operator are the device ones (``localization.matchers.*``,
``pose_estimator.absolute_pose_estimation`` instead of pycolmap).  The OpenCV windows of the reference's ``run``
(``config['localization']['show']``) belong to its viewer and are not reproduced.
Tracking is sequential per sequence, so multi-GPU sharding is per sequence in this mode."""
from __future__ import annotations

import time
from typing import Callable, Optional

import numpy as np
import torch

from . import pose_estimator


class Tracker:
    def __init__(self, locMap=None, matcher=None, config: Optional[dict] = None, device='cuda', pose_fn: Optional[Callable] = None):
        self.locMap, self.matcher, self.config = locMap, matcher, config or {}
        self.loc_config = self.config.get('localization', {})
        self.device = torch.device(device)
        self.pose_fn = pose_fn or pose_estimator.absolute_pose_estimation
        self.lost = True
        self.curr_frame = None
        self.last_frame = None

    # -- matcher call shared by the tracking variants --------------------------------------------------------------
    def _match(self, kpts0, scores0, descs0, cam0, kpts1, scores1, descs1, cam1) -> np.ndarray:
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))[None].float().to(self.device)
        with torch.no_grad():
            out = self.matcher({
                'descriptors0': t(descs0), 'keypoints0': t(kpts0), 'scores0': t(scores0),
                'image_shape0': (1, 3, cam0.width, cam0.height),   # the reference's (W, H) order, kept
                'descriptors1': t(descs1), 'keypoints1': t(kpts1), 'scores1': t(scores1),
                'image_shape1': (1, 3, cam1.width, cam1.height),
            })
        return out['matches0'][0].cpu().numpy()

    def _pose(self, ret_extra: dict, p2d: np.ndarray, p3d: np.ndarray, camera) -> dict:
        ret = self.pose_fn(p2d + 0.5, p3d, camera,
                           estimation_options={'ransac': {'max_error': self.config['localization']['threshold']}},
                           refinement_options={})
        if ret is None:
            ret = {'success': False}
        else:
            ret['success'] = True
            ret['qvec'] = np.asarray(ret['cam_from_world'].rotation.quat)[[3, 0, 1, 2]]
            ret['tvec'] = ret['cam_from_world'].translation
        ret.update(ret_extra)
        return ret

    def track_last_frame(self, curr_frame, last_frame) -> dict:
        curr_kpts, last_kpts = curr_frame.keypoints[:, :2], last_frame.keypoints[:, :2]
        indices = self._match(curr_kpts, curr_frame.keypoints[:, 2], curr_frame.descriptors, curr_frame.camera,
                              last_kpts, last_frame.keypoints[:, 2], last_frame.descriptors, last_frame.camera)
        q_ids = np.nonzero(indices >= 0)[0]
        ref = indices[q_ids]
        pids = np.asarray(last_frame.point3D_ids)[ref]
        keep = pids >= 0                      # keypoints of the last frame without a 3-D point carry id -1
        q_ids, ref, pids = q_ids[keep], ref[keep], pids[keep]
        matched_kpts, matched_xyzs = curr_kpts[q_ids], np.asarray(last_frame.xyzs)[ref]
        print('Tracking: {:d} matches from {:d}-{:d} kpts'.format(matched_kpts.shape[0], curr_kpts.shape[0], last_kpts.shape[0]))
        return self._pose({'matched_keypoints': matched_kpts, 'matched_keypoint_ids': q_ids, 'matched_ref_keypoints': last_kpts[ref],
                           'matched_xyzs': matched_xyzs, 'matched_point3D_ids': pids,
                           'matched_sids': np.asarray(last_frame.seg_ids)[ref],
                           'reference_frame_id': last_frame.reference_frame_id,
                           'matched_scene_name': last_frame.matched_scene_name}, matched_kpts, matched_xyzs, curr_frame.camera)

    def track_last_frame_fast(self, curr_frame, last_frame) -> dict:
        """reference :235-311: only the last frame's keypoints WITH a 3-D point take part, and only the current keypoints
        inside their bounding box.  (The reference still uses pycolmap's pre-0.5 call signature here and does not handle
        a failed estimate; the result convention of ``track_last_frame`` is used instead.)"""
        has3d = np.asarray(last_frame.point3D_ids) >= 0
        last_kpts = last_frame.keypoints[:, :2][has3d]
        curr_kpts = curr_frame.keypoints[:, :2]
        box = (curr_kpts[:, 0] >= last_kpts[:, 0].min()) & (curr_kpts[:, 0] <= last_kpts[:, 0].max()) & \
              (curr_kpts[:, 1] >= last_kpts[:, 1].min()) & (curr_kpts[:, 1] <= last_kpts[:, 1].max())
        curr_ids = np.nonzero(box)[0]
        curr_kpts = curr_kpts[box]
        indices = self._match(curr_kpts, curr_frame.keypoints[:, 2][box], curr_frame.descriptors[box], curr_frame.camera,
                              last_kpts, last_frame.keypoints[:, 2][has3d], last_frame.descriptors[has3d], last_frame.camera)
        valid = indices >= 0
        ref = indices[valid]
        matched_kpts, matched_xyzs = curr_kpts[valid], np.asarray(last_frame.xyzs)[has3d][ref]
        print('Tracking: {:d} matches from {:d}-{:d} kpts'.format(matched_kpts.shape[0], curr_kpts.shape[0], last_kpts.shape[0]))
        return self._pose({'matched_keypoints': matched_kpts, 'matched_keypoint_ids': curr_ids[valid],
                           'matched_ref_keypoints': last_kpts[ref], 'matched_xyzs': matched_xyzs,
                           'matched_point3D_ids': np.asarray(last_frame.point3D_ids)[has3d][ref],
                           'matched_sids': np.asarray(last_frame.seg_ids)[has3d][ref],
                           'reference_frame_id': last_frame.reference_frame_id,
                           'matched_scene_name': last_frame.matched_scene_name}, matched_kpts, matched_xyzs, curr_frame.camera)

    @torch.no_grad()
    def match_frame(self, frame, reference_frame):
        """reference :313-338: matched keypoint index pairs (frame, reference) whose reference keypoint has a 3-D point.
        ``image_size`` is (height, width) and is passed in that order, as the reference does."""
        cam = lambda f: type('C', (), {'width': f.image_size[0], 'height': f.image_size[1]})
        m = self._match(frame.keypoints[:, :2], frame.keypoints[:, 2], frame.descriptors, cam(frame),
                        reference_frame.keypoints[:, :2], reference_frame.keypoints[:, 2], reference_frame.descriptors,
                        cam(reference_frame))
        ids1 = np.nonzero(m >= 0)[0]
        ids2 = m[ids1]
        ok = np.asarray(reference_frame.points3d_mask)[ids2]
        return ids1[ok], ids2[ok]

    # -- the tracking step of the online loop (reference :37-160, without the OpenCV windows) ------------------------------
    def verify_and_update(self, q_frame, ret: dict) -> bool:
        q_frame.qvec, q_frame.tvec = ret['qvec'], ret['tvec']
        q_err, t_err = q_frame.compute_pose_error()
        if ret['num_inliers'] < self.loc_config['min_inliers']:
            print('Failed due to insufficient {:d} inliers,  q_err: {:.2f}, t_err: {:.2f}'.format(ret['num_inliers'], q_err, t_err))
            q_frame.tracking_status = False
            q_frame.clear_localization_track()
            return False
        print('Succeed! Find {}/{} 2D-3D inliers,q_err: {:.2f}, t_err: {:.2f}'.format(ret['num_inliers'], ret['matched_keypoints'].shape[0],
                                                                                   q_err, t_err))
        q_frame.tracking_status = True
        self.update_current_frame(curr_frame=q_frame, ret=ret)
        return True

    def update_current_frame(self, curr_frame, ret: dict):
        curr_frame.qvec, curr_frame.tvec = ret['qvec'], ret['tvec']
        curr_frame.matched_scene_name = ret['matched_scene_name']
        curr_frame.reference_frame_id = ret['reference_frame_id']
        inl = np.array(ret['inliers'])
        for key in ('matched_keypoints', 'matched_xyzs', 'matched_point3D_ids', 'matched_keypoint_ids', 'matched_sids'):
            setattr(curr_frame, key, ret[key][inl])

    def run(self, frame) -> bool:
        print('Start tracking...')
        self.curr_frame = frame
        t0 = time.time()
        ret = self.track_last_frame(curr_frame=frame, last_frame=self.last_frame)
        frame.time_loc = frame.time_loc + time.time() - t0
        if not ret['success']:
            return False
        ret['matched_scene_name'] = self.last_frame.scene_name
        if not self.verify_and_update(q_frame=frame, ret=ret):
            return False
        success = True
        if ret['num_inliers'] < 256:  # refinement is necessary for tracking last frame
            t0 = time.time()
            ret = self.locMap.sub_maps[self.last_frame.matched_scene_name].refine_pose(
                frame, refinement_method=self.loc_config['refinement_method'])
            frame.time_ref = frame.time_ref + time.time() - t0
            ret['matched_scene_name'] = self.last_frame.scene_name
            success = self.verify_and_update(q_frame=frame, ret=ret)
        self.lost = success
        return success
