"""Device-resident reference frames and the two hot functions of the reference's ``SingleMap3D``.

Scope: ``SingleMap3D.localize_with_ref_frame`` and ``SingleMap3D.match``
(reference ``localization/singlemap3d.py:127-226``) and the ``RefFrame`` accessors they read
(``localization/refframe.py:34-75``).  Same arguments, same result dicts (numpy, same keys) -- but a reference frame's
keypoints / descriptors / scores are uploaded ONCE per (frame, landmark) and stay resident (the reference re-uploads
them on every matcher call, ``singlemap3d.py:143-153``), and the pose comes from the GPU P3P / RANSAC operator instead
of ``pycolmap.absolute_pose_estimation`` (``singlemap3d.py:168-175``).  The id bookkeeping is host numpy, like the reference's.  Map loading (COLMAP
models, ``point3D_desc.npy`` ...) stays outside (``map_io.load_single_map``): the constructor takes already-associated
reference frames.  The refinement entry points ``refine_pose`` / ``refine_pose_by_matching`` / ``refine_pose_by_projection``
(reference :260-498), ``build_covisibility_graph`` / ``find_reference_frames`` / ``check_semantic_consistency`` are here too.
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
                 point3D_sids: Dict[int, int], device='cuda', pose_fn: Optional[Callable] = None, point3Ds: Optional[dict] = None,
                 start_sid: int = 0):
        """``config['localization']['threshold']`` = RANSAC max_error in pixels (configs/*.yaml: 8 or 12);
        ``seg_ref_frame_ids[sid]`` = virtual reference frame ids of landmark ``sid`` (first = best);
        ``point3D_sids[point3D_id]`` = segment id of a 3-D point (``self.point3Ds[v].seg_id`` in the reference);
        ``pose_fn`` defaults to the GPU ``absolute_pose_estimation`` (same signature as pycolmap's);
        ``point3Ds[id]`` = objects with ``xyz``, ``descriptor``, ``seg_id``, ``frame_ids`` (reference point3d.py) -- needed by
        the refinement entry points only; they are packed once into device-resident tables."""
        self.config, self.matcher = config, matcher
        self.reference_frames, self.seg_ref_frame_ids, self.point3D_sids = reference_frames, seg_ref_frame_ids, point3D_sids
        self.device = torch.device(device)
        self.pose_fn = pose_fn or pose_estimator.absolute_pose_estimation
        self.start_sid = start_sid
        self.point3Ds = point3Ds
        self.covisible_graph: Dict[int, list] = {}
        self._p3d_table = None
        if point3Ds is not None and 'covisibility_frame' in config.get('localization', {}):
            vrf = [f for ids in seg_ref_frame_ids.values() for f in ids]
            self.build_covisibility_graph([v for v in np.unique(vrf) if v in reference_frames],
                                          n_frame=config['localization']['covisibility_frame'])

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

    # -- covisibility (reference singlemap3d.py:228-259, 500-513) -----------------------------------------------------
    def build_covisibility_graph(self, frame_ids=None, n_frame: int = 20):
        """For every frame in ``frame_ids``: the frames sharing most 3-D points with it (itself included, like the
        reference), at most ``n_frame``, most shared first."""
        from collections import defaultdict
        frame_ids = list(self.reference_frames.keys()) if frame_ids is None else frame_ids
        self.covisible_graph = {}
        for fid in frame_ids:
            votes = defaultdict(int)
            for pid in self.reference_frames[fid].point3D_ids:
                if pid == -1 or pid not in self.point3Ds:
                    continue
                for img_id in self.point3Ds[pid].frame_ids:
                    votes[img_id] += 1
            ids = np.array(list(votes.keys()))
            num = np.array([votes[i] for i in ids])
            if len(ids) <= n_frame:
                self.covisible_graph[fid] = ids[np.argsort(-num)]
            else:
                top = np.argpartition(num, -n_frame)[-n_frame:]
                self.covisible_graph[fid] = [ids[i] for i in top[np.argsort(-num[top])]]

    def find_reference_frames(self, matched_point3D_ids, candidate_frame_ids=None):
        """Candidate frames ordered by how many of the matched 3-D points they observe (reference :500-513)."""
        from collections import defaultdict
        votes = defaultdict(int)
        for pid in matched_point3D_ids:
            for im_id in self.point3Ds[pid].frame_ids:
                if candidate_frame_ids is not None and im_id in candidate_frame_ids:
                    votes[im_id] += 1
        ids = np.array(list(votes.keys()))
        num = np.array([votes[i] for i in ids])
        return ids[np.argsort(num)[::-1]]

    def check_semantic_consistency(self, q_frame, sid, overlap_ratio: float = 0.5) -> bool:
        """reference :515-532: the query's and the reference frame's landmark ids must overlap on both sides."""
        ref_frame = self.reference_frames[self.seg_ref_frame_ids[sid][0]]
        q_sids = q_frame.seg_ids
        ref_sids = np.array([self.point3D_sids[v] for v in ref_frame.point3D_ids]) + self.start_sid
        common = np.intersect1d(q_sids, ref_sids)
        n1 = sum(int(np.sum(q_sids == s)) for s in common)
        n2 = sum(int(np.sum(ref_sids == s)) for s in common)
        return min(n1 / q_sids.shape[0], n2 / ref_sids.shape[0]) >= overlap_ratio

    # -- refinement (reference singlemap3d.py:260-498) ------------------------------------------------------------------
    def refine_pose(self, q_frame, refinement_method: str = 'matching'):
        if refinement_method == 'matching':
            return self.refine_pose_by_matching(q_frame=q_frame)
        if refinement_method == 'projection':
            return self.refine_pose_by_projection(q_frame=q_frame)
        raise NotImplementedError

    def _finish_refinement(self, ret: Optional[dict], point3D_ids: np.ndarray) -> dict:
        if ret is None:
            ret = {'success': False}
        else:
            ret['success'] = True
            ret['qvec'] = np.asarray(ret['cam_from_world'].rotation.quat)[[3, 0, 1, 2]]
            ret['tvec'] = ret['cam_from_world'].translation
        voters = point3D_ids[np.array(ret['inliers'])] if ret['success'] else point3D_ids
        best = self.find_reference_frames(matched_point3D_ids=voters, candidate_frame_ids=self.covisible_graph.keys())
        ret['refinement_reference_frame_ids'] = best[:self.config['localization']['covisibility_frame']]
        ret['reference_frame_id'] = best[0]
        return ret

    def refine_pose_by_matching(self, q_frame) -> dict:
        """reference :268-365: match the query against every frame covisible with its reference frame (one matcher call
        each, reference features resident on the device), pool the 2D-3D matches with the ones of the first localisation
        and estimate ONE pose with ``min_num_trials`` 1000 / ``max_num_trials`` 10000 / ``confidence`` 0.995."""
        ref_frame_id = q_frame.reference_frame_id
        db_ids = self.covisible_graph[ref_frame_id]
        print('Find {} covisible frames'.format(len(db_ids)))
        init = None
        if q_frame.tracking_status and ref_frame_id in db_ids:
            ids0 = np.asarray(q_frame.matched_point3D_ids)
            init = (q_frame.matched_keypoints, np.array([self.point3Ds[v].xyz for v in ids0]).reshape(-1, 3), ids0,
                    q_frame.matched_keypoint_ids)
        query = {'keypoints': q_frame.keypoints[:, :2], 'scores': q_frame.keypoints[:, 2], 'descriptors': q_frame.descriptors,
                 'camera': q_frame.camera}
        parts = []
        for frame_id in db_ids:
            out = self.match(query_data=query, ref_data=self.reference_frames[frame_id].get_keypoints())
            if out['matched_keypoints'].shape[0] > 0:
                parts.append((out['matched_keypoints'], out['matched_xyzs'], out['matched_point3D_ids'], out['matched_keypoint_ids']))
        if init is not None and init[0].shape[0] > 0:
            parts.append(init)
        if not parts:
            raise ValueError('refine_pose_by_matching: no 2D-3D match in any covisible frame '
                             '(the reference fails with an IndexError here, singlemap3d.py:312)')
        kpts = np.vstack([p[0] for p in parts])
        xyzs = np.vstack([p[1] for p in parts]).reshape(-1, 3)
        pids = np.hstack([p[2] for p in parts])
        kids = np.hstack([p[3] for p in parts])
        print('Refinement by matching. Get {:d} covisible frames with {:d} matches for optimization'.format(len(db_ids), xyzs.shape[0]))
        ret = self.pose_fn(kpts + 0.5, xyzs, q_frame.camera,
                           estimation_options={'ransac': {'max_error': self.config['localization']['threshold'],
                                                          'min_num_trials': 1000, 'max_num_trials': 10000, 'confidence': 0.995}},
                           refinement_options={})
        ret = dict(ret) if ret is not None else None
        sids = np.array([self.point3Ds[v].seg_id for v in pids])
        out = self._finish_refinement(ret, pids)
        out.update({'matched_keypoints': kpts, 'matched_keypoint_ids': kids, 'matched_xyzs': xyzs, 'matched_point3D_ids': pids,
                    'matched_sids': sids})
        return out

    def _point_table(self):
        """All labelled 3-D points as device-resident tables (built once): sorted ids, xyz, descriptors, segment ids."""
        if self._p3d_table is None:
            ids = np.array(sorted(self.point3Ds.keys()), dtype=np.int64)
            xyz = np.array([self.point3Ds[i].xyz for i in ids], dtype=np.float64).reshape(-1, 3)
            desc = np.array([self.point3Ds[i].descriptor for i in ids], dtype=np.float32)
            sid = np.array([self.point3Ds[i].seg_id for i in ids])
            self._p3d_table = {'ids': ids, 'xyz': xyz, 'sid': sid,
                               'xyz_dev': torch.from_numpy(xyz).float().to(self.device),
                               'desc_dev': torch.from_numpy(desc).to(self.device)}
        return self._p3d_table

    @torch.no_grad()
    def refine_pose_by_projection(self, q_frame) -> dict:
        """reference :367-498 (K18, the default ``refinement_method``): project the 3-D points of the frames covisible with
        the query's reference frame with the current pose, pair every query keypoint with the visible point of smallest
        descriptor distance inside a ``2 * threshold`` pixel window (ratio test 0.995), estimate the pose from those pairs.
        Projection, the M x N similarity (tcgen05 GEMM) and the masked top-2 run on the device against the resident point
        tables; only the pose and the matched indices come back."""
        from .. import ops
        from .map_io import qvec2rotmat
        cam = q_frame.camera
        fx, fy, cx, cy = pose_estimator.camera_intrinsics(cam)   # distortion is not included, as in the reference (:399)
        ref_id = q_frame.reference_frame_id
        covis = self.covisible_graph[ref_id]
        if ref_id not in covis:
            covis.append(ref_id)
        tab = self._point_table()
        all_ids = np.unique(np.concatenate([np.asarray(self.reference_frames[f].point3D_ids).ravel() for f in covis]))
        rows = np.searchsorted(tab['ids'], all_ids)
        rows_dev = torch.from_numpy(rows).to(self.device)
        q_kpts = torch.from_numpy(np.ascontiguousarray(q_frame.keypoints[:, :2])).float().to(self.device)
        q_desc = torch.from_numpy(np.ascontiguousarray(q_frame.descriptors)).float().to(self.device)
        match, _, _ = ops.match_by_projection(q_kpts, q_desc, tab['xyz_dev'][rows_dev].contiguous(), tab['desc_dev'][rows_dev].contiguous(),
                                              qvec2rotmat(q_frame.qvec), np.asarray(q_frame.tvec, float), fx, fy, cx, cy, cam.width,
                                              cam.height, self.config['localization']['threshold'])
        match = match.cpu().numpy()
        keep = match >= 0
        sel = rows[match[keep]]
        mkpts, mkpt_ids = q_frame.keypoints[keep], np.where(keep)[0]
        mxyzs, mpids, msids = tab['xyz'][sel], tab['ids'][sel], tab['sid'][sel]
        ret = self.pose_fn(mkpts[:, :2] + 0.5, mxyzs, cam,
                           estimation_options={'ransac': {'max_error': self.config['localization']['threshold']}},
                           refinement_options={})
        ret = dict(ret) if ret is not None else None
        out = self._finish_refinement(ret, mpids)
        print('Refinement by projection. Get {:d} inliers of {:d} matches for optimization'.format(
            out['num_inliers'] if out['success'] else 0, mxyzs.shape[0]))
        out.update({'matched_keypoints': mkpts, 'matched_xyzs': mxyzs, 'matched_point3D_ids': mpids, 'matched_sids': msids,
                    'matched_keypoint_ids': mkpt_ids})
        if not out['success']:
            out['num_inliers'] = 0
            out['inliers'] = np.zeros(shape=(mkpts.shape[0],), dtype=bool)
        return out
