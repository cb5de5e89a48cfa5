"""On-disk map formats: COLMAP binary models, PRAM's compressed models and the landmark
``.npy`` dictionaries, and the loader that turns a PRAM landmark folder into device-resident reference frames.

Formats:
* ``cameras.bin`` / ``images.bin`` / ``points3D.bin``: COLMAP's public binary layout (``src/base/reconstruction.cc``;
  the reference reads them with ``colmap_utils/read_write_model.py:127-407``).
* compressed model (reference ``read_write_model.py:433-554``): ``images.bin`` keeps only the ``point3D_id`` of every
  2-D point (int64, no coordinates), ``points3D.bin`` only the image ids of every track (int32, no point2D index).
* ``point3D_desc.npy`` (dict id -> descriptor), ``point3D_cluster_n{K}_{mode}_{method}.npy`` (``{'id', 'label'}``),
  ``point3D_vrf_n{K}_{mode}_{method}.npy`` (dict sid -> {vi: {'image_id', 'original_points3d', ...}}): pickled numpy
  objects, read exactly like the reference (``localization/singlemap3d.py:30-66``).

The readers are written against the format, not the reference's code: records are decoded with ``struct`` /
``numpy.frombuffer`` from one ``bytes`` object per file.
"""
from __future__ import annotations

import os.path as osp
import struct
from collections import namedtuple
from typing import Dict, Tuple

import numpy as np

Camera = namedtuple('Camera', ['id', 'model', 'width', 'height', 'params'])
Image = namedtuple('Image', ['id', 'qvec', 'tvec', 'camera_id', 'name', 'xys', 'point3D_ids'])
Point3D = namedtuple('Point3D', ['id', 'xyz', 'rgb', 'error', 'image_ids', 'point2D_idxs'])

# COLMAP camera model id -> (name, number of parameters)
CAMERA_MODELS = {0: ('SIMPLE_PINHOLE', 3), 1: ('PINHOLE', 4), 2: ('SIMPLE_RADIAL', 4), 3: ('RADIAL', 5), 4: ('OPENCV', 8),
                 5: ('OPENCV_FISHEYE', 8), 6: ('FULL_OPENCV', 12), 7: ('FOV', 5), 8: ('SIMPLE_RADIAL_FISHEYE', 4),
                 9: ('RADIAL_FISHEYE', 5), 10: ('THIN_PRISM_FISHEYE', 12)}


def read_cameras_binary(path: str) -> Dict[int, Camera]:
    buf = open(path, 'rb').read()
    (n,), off = struct.unpack_from('<Q', buf, 0), 8
    cams = {}
    for _ in range(n):
        cid, model_id, width, height = struct.unpack_from('<iiQQ', buf, off)
        off += 24
        name, npar = CAMERA_MODELS[model_id]
        params = np.frombuffer(buf, '<f8', npar, off).copy()
        off += 8 * npar
        cams[cid] = Camera(id=cid, model=name, width=width, height=height, params=params)
    return cams


def _read_images(path: str, compressed: bool) -> Dict[int, Image]:
    buf = open(path, 'rb').read()
    (n,), off = struct.unpack_from('<Q', buf, 0), 8
    images = {}
    for _ in range(n):
        rec = struct.unpack_from('<idddddddi', buf, off)
        off += 64
        end = buf.index(b'\x00', off)
        name = buf[off:end].decode('utf-8')
        off = end + 1
        (m,) = struct.unpack_from('<Q', buf, off)
        off += 8
        if compressed:
            xys = np.array([])
            ids = np.frombuffer(buf, '<i8', m, off).copy()
            off += 8 * m
        else:
            pts = np.frombuffer(buf, np.dtype([('x', '<f8'), ('y', '<f8'), ('id', '<i8')]), m, off)
            xys = np.column_stack([pts['x'], pts['y']]) if m else np.zeros((0, 2))
            ids = pts['id'].copy()
            off += 24 * m
        images[rec[0]] = Image(id=rec[0], qvec=np.array(rec[1:5]), tvec=np.array(rec[5:8]), camera_id=rec[8], name=name,
                               xys=xys, point3D_ids=ids)
    return images


def _read_points3d(path: str, compressed: bool) -> Dict[int, Point3D]:
    buf = open(path, 'rb').read()
    (n,), off = struct.unpack_from('<Q', buf, 0), 8
    pts = {}
    for _ in range(n):
        pid, x, y, z, r, g, b, err = struct.unpack_from('<QdddBBBd', buf, off)
        off += 43
        (t,) = struct.unpack_from('<Q', buf, off)
        off += 8
        if compressed:
            image_ids = np.frombuffer(buf, '<i4', t, off).copy()
            idxs = np.array([])
            off += 4 * t
        else:
            tr = np.frombuffer(buf, '<i4', 2 * t, off)
            image_ids, idxs = tr[0::2].copy(), tr[1::2].copy()
            off += 8 * t
        pts[pid] = Point3D(id=pid, xyz=np.array([x, y, z]), rgb=np.array([r, g, b]), error=np.array(err), image_ids=image_ids,
                           point2D_idxs=idxs)
    return pts


def read_images_binary(path): return _read_images(path, False)
def read_points3D_binary(path): return _read_points3d(path, False)
def read_compressed_images_binary(path): return _read_images(path, True)
def read_compressed_points3d_binary(path): return _read_points3d(path, True)


def read_model(path: str, ext: str = '.bin') -> Tuple[dict, dict, dict]:
    assert ext == '.bin', 'only the binary model format is supported'
    return (read_cameras_binary(osp.join(path, 'cameras.bin')), read_images_binary(osp.join(path, 'images.bin')),
            read_points3D_binary(osp.join(path, 'points3D.bin')))


def read_compressed_model(path: str, ext: str = '.bin') -> Tuple[dict, dict, dict]:
    assert ext == '.bin', 'only the binary model format is supported'
    return (read_cameras_binary(osp.join(path, 'cameras.bin')), read_compressed_images_binary(osp.join(path, 'images.bin')),
            read_compressed_points3d_binary(osp.join(path, 'points3D.bin')))


def qvec2rotmat(q) -> np.ndarray:
    w, x, y, z = q
    return np.array([[1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * z * x + 2 * w * y],
                     [2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x],
                     [2 * z * x - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x * x - 2 * y * y]])


def project(camera: Camera, qvec, tvec, xyzs: np.ndarray) -> np.ndarray:
    """Pinhole projection without distortion, like reference RefFrame.project (refframe.py:131-147)."""
    from .pose_estimator import camera_intrinsics
    fx, fy, cx, cy = camera_intrinsics(camera)
    xc = xyzs @ qvec2rotmat(qvec).T + np.asarray(tvec, float).reshape(1, 3)
    return np.stack([fx * xc[:, 0] / xc[:, 2] + cx, fy * xc[:, 1] / xc[:, 2] + cy], 1)


def load_single_map(config: dict, matcher, with_compress: bool = True, device='cuda', pose_fn=None):
    """Build a ``SingleMap3D`` from a PRAM landmark folder -- the data side of the reference constructor
    (``localization/singlemap3d.py:25-97``): model + descriptor / cluster / virtual-reference-frame dictionaries ->
    per-frame keypoints (projections of the frame's 3-D points), descriptors, xyz, ids and segment ids
    (``RefFrame.associate_keypoints_with_point3Ds``, refframe.py:99-129), resident on ``device``.
    Frames without any labelled 3-D point are dropped, as in the reference.  The covisibility graph is not built."""
    from .singlemap3d import RefFrame, SingleMap3D
    root = config['landmark_path']
    tag = 'n{:d}_{:s}_{:s}'.format(config['n_cluster'], config['cluster_mode'], config['cluster_method'])
    if with_compress:
        mdir = osp.join(root, 'compress_model_{:s}'.format(config['cluster_method']))
        cameras, images, p3ds = read_compressed_model(mdir)
    else:
        mdir = osp.join(root, 'model')
        cameras, images, p3ds = read_model(mdir)
    desc_path = osp.join(mdir, 'point3D_desc.npy') if with_compress else osp.join(root, 'point3D_desc.npy')
    p3d_descs = np.load(desc_path, allow_pickle=True)[()]
    seg_data = np.load(osp.join(root, f'point3D_cluster_{tag}.npy'), allow_pickle=True)[()]
    p3d_seg = {int(i): int(s) for i, s in zip(seg_data['id'], seg_data['label'])}
    seg_vrf = np.load(osp.join(root, f'point3D_vrf_{tag}.npy'), allow_pickle=True)[()]
    # 3-D point table restricted to labelled points (reference initialize_point3Ds)
    points = {pid: p for pid, p in p3ds.items() if pid in p3d_seg}
    frame_p3d_ids = {fid: np.asarray(im.point3D_ids) for fid, im in images.items()}
    seg_ref_frame_ids = {}
    for sid, vrfs in seg_vrf.items():
        seg_ref_frame_ids[sid] = []
        for vi in vrfs.keys():
            fid = vrfs[vi]['image_id']
            seg_ref_frame_ids[sid].append(fid)
            if with_compress and fid in frame_p3d_ids:
                frame_p3d_ids[fid] = np.asarray(vrfs[vi]['original_points3d'])
    frames = {}
    for fid, im in images.items():
        ids = np.array([v for v in frame_p3d_ids[fid] if v in points], dtype=np.int64)
        if ids.size == 0:
            continue
        xyzs = np.array([points[v].xyz for v in ids])
        descs = np.array([p3d_descs[v] for v in ids])
        scores = 1 / np.clip(np.array([float(points[v].error) for v in ids]) * 5, a_min=1., a_max=20.)
        cam = cameras[im.camera_id]
        uvs = project(cam, im.qvec, im.tvec, xyzs)
        frames[fid] = RefFrame(camera=cam, id=fid, keypoints=np.hstack([uvs, scores.reshape(-1, 1)]), descriptors=descs, xyzs=xyzs,
                               point3D_ids=ids, keypoint_segs=np.array([p3d_seg[int(v)] for v in ids]), qvec=im.qvec, tvec=im.tvec,
                               name=im.name, device=device)
    return SingleMap3D(config, matcher, frames, seg_ref_frame_ids, p3d_seg, device=device, pose_fn=pose_fn)
