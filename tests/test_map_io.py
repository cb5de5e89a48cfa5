"""On-disk map formats (CPU): COLMAP binary + PRAM compressed models written by a small struct writer (and, when the
reference tree is mounted, by the reference's own writers / read back by its own readers), then the landmark-folder
loader that builds device-resident reference frames."""
import struct
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import ref_loader as RL
from pram_b200.localization import map_io as M


def _synthetic(rs, n_img=4, n_pts=60):
    cams = {1: M.Camera(1, 'PINHOLE', 640, 480, np.array([525.0, 520.0, 320.0, 240.0])),
            2: M.Camera(2, 'SIMPLE_RADIAL', 800, 600, np.array([700.0, 400.0, 300.0, 0.01]))}
    pts = {}
    for i in range(n_pts):
        pid = 1000 + 3 * i
        t = rs.randint(2, 5)
        pts[pid] = M.Point3D(pid, rs.randn(3) + np.array([0, 0, 5.0]), rs.randint(0, 256, 3), np.array(rs.rand() * 2),
                             rs.randint(1, n_img + 1, t).astype(np.int32), rs.randint(0, 50, t).astype(np.int32))
    imgs = {}
    keys = list(pts)
    for i in range(1, n_img + 1):
        m = rs.randint(10, 30)
        ids = rs.choice(keys + [-1] * 10, m).astype(np.int64)
        q = rs.randn(4); q /= np.linalg.norm(q)
        imgs[i] = M.Image(i, q, rs.randn(3) * 0.1, 1 + i % 2, f'seq/frame_{i:03d}.png', rs.rand(m, 2) * 400, ids)
    return cams, imgs, pts


def _write(path, cams, imgs, pts, compressed):
    inv = {v[0]: k for k, v in M.CAMERA_MODELS.items()}
    with open(path / 'cameras.bin', 'wb') as f:
        f.write(struct.pack('<Q', len(cams)))
        for c in cams.values():
            f.write(struct.pack('<iiQQ', c.id, inv[c.model], c.width, c.height) + np.asarray(c.params, '<f8').tobytes())
    with open(path / 'images.bin', 'wb') as f:
        f.write(struct.pack('<Q', len(imgs)))
        for im in imgs.values():
            f.write(struct.pack('<idddddddi', im.id, *im.qvec, *im.tvec, im.camera_id) + im.name.encode() + b'\x00')
            f.write(struct.pack('<Q', len(im.point3D_ids)))
            for j, pid in enumerate(im.point3D_ids):
                f.write(struct.pack('<q', pid) if compressed else struct.pack('<ddq', *im.xys[j], pid))
    with open(path / 'points3D.bin', 'wb') as f:
        f.write(struct.pack('<Q', len(pts)))
        for p in pts.values():
            f.write(struct.pack('<QdddBBBd', p.id, *p.xyz, *[int(v) for v in p.rgb], float(p.error)))
            f.write(struct.pack('<Q', len(p.image_ids)))
            for j, iid in enumerate(p.image_ids):
                f.write(struct.pack('<i', iid) if compressed else struct.pack('<ii', iid, p.point2D_idxs[j]))


@pytest.mark.parametrize('compressed', [False, True])
def test_model_round_trip(tmp_path, compressed):
    cams, imgs, pts = _synthetic(np.random.RandomState(0))
    _write(tmp_path, cams, imgs, pts, compressed)
    c2, i2, p2 = (M.read_compressed_model if compressed else M.read_model)(str(tmp_path))
    assert set(c2) == set(cams) and set(i2) == set(imgs) and set(p2) == set(pts)
    for k, c in cams.items():
        assert (c2[k].model, c2[k].width, c2[k].height) == (c.model, c.width, c.height) and np.array_equal(c2[k].params, c.params)
    for k, im in imgs.items():
        assert i2[k].name == im.name and i2[k].camera_id == im.camera_id
        assert np.array_equal(i2[k].qvec, im.qvec) and np.array_equal(i2[k].tvec, im.tvec)
        assert np.array_equal(i2[k].point3D_ids, im.point3D_ids)
        assert i2[k].xys.size == 0 if compressed else np.array_equal(i2[k].xys, im.xys)
    for k, p in pts.items():
        assert np.array_equal(p2[k].xyz, p.xyz) and np.array_equal(p2[k].rgb, p.rgb) and float(p2[k].error) == float(p.error)
        assert np.array_equal(p2[k].image_ids, p.image_ids)
        assert p2[k].point2D_idxs.size == 0 if compressed else np.array_equal(p2[k].point2D_idxs, p.point2D_idxs)


@pytest.mark.skipif(not RL.reference_available(), reason='reference tree not mounted')
@pytest.mark.parametrize('compressed', [False, True])
def test_readers_agree_with_reference_readers(tmp_path, compressed):
    """Files written by the reference's own writers are decoded identically by our readers and by the reference's."""
    RL.import_reference()
    from colmap_utils import read_write_model as RW
    cams, imgs, pts = _synthetic(np.random.RandomState(1))
    rc = {k: RW.Camera(id=c.id, model=c.model, width=c.width, height=c.height, params=c.params) for k, c in cams.items()}
    ri = {k: RW.Image(id=i.id, qvec=i.qvec, tvec=i.tvec, camera_id=i.camera_id, name=i.name, xys=i.xys, point3D_ids=i.point3D_ids)
          for k, i in imgs.items()}
    rp = {k: RW.Point3D(id=p.id, xyz=p.xyz, rgb=p.rgb, error=float(p.error), image_ids=p.image_ids, point2D_idxs=p.point2D_idxs)
          for k, p in pts.items()}
    RW.write_cameras_binary(rc, str(tmp_path / 'cameras.bin'))
    if compressed:
        RW.write_compressed_images_binary(ri, str(tmp_path / 'images.bin'))
        RW.write_compressed_points3d_binary(rp, str(tmp_path / 'points3D.bin'))
        ref = RW.read_compressed_model(str(tmp_path), '.bin')
        ours = M.read_compressed_model(str(tmp_path))
    else:
        RW.write_images_binary(ri, str(tmp_path / 'images.bin'))
        RW.write_points3d_binary(rp, str(tmp_path / 'points3D.bin'))
        ref = RW.read_model(str(tmp_path), '.bin')
        ours = M.read_model(str(tmp_path))
    for a, b in zip(ref, ours):
        assert set(a) == set(b)
    for k in ref[0]:
        assert ref[0][k].model == ours[0][k].model and np.array_equal(ref[0][k].params, ours[0][k].params)
    for k in ref[1]:
        assert ref[1][k].name == ours[1][k].name and np.array_equal(ref[1][k].qvec, ours[1][k].qvec)
        assert np.array_equal(ref[1][k].point3D_ids, ours[1][k].point3D_ids) and np.array_equal(ref[1][k].xys, ours[1][k].xys)
    for k in ref[2]:
        assert np.array_equal(ref[2][k].xyz, ours[2][k].xyz) and np.array_equal(ref[2][k].image_ids, ours[2][k].image_ids)
        assert np.array_equal(ref[2][k].rgb, ours[2][k].rgb) and float(ref[2][k].error) == float(ours[2][k].error)


def test_load_single_map(tmp_path):
    """Landmark folder -> SingleMap3D with resident reference frames: keypoints are the projections of the frame's
    labelled 3-D points, scores follow 1 / clip(5 error, 1, 20), virtual reference frames take their point list from
    the vrf dictionary (compressed maps), unlabelled points and empty frames are dropped."""
    rs = np.random.RandomState(2)
    cams, imgs, pts = _synthetic(rs, n_img=5, n_pts=80)
    mdir = tmp_path / 'compress_model_birch'
    mdir.mkdir()
    _write(mdir, cams, imgs, pts, compressed=True)
    keys = np.array(list(pts))
    labelled = keys[rs.rand(keys.size) < 0.8]
    np.save(mdir / 'point3D_desc.npy', {int(k): rs.randn(128).astype(np.float32) for k in keys}, allow_pickle=True)
    np.save(tmp_path / 'point3D_cluster_n8_xz_birch.npy', {'id': labelled, 'label': rs.randint(1, 9, labelled.size)}, allow_pickle=True)
    vrf_pts = rs.choice(labelled, 25, replace=False)
    np.save(tmp_path / 'point3D_vrf_n8_xz_birch.npy', {3: {0: {'image_id': 2, 'original_points3d': vrf_pts}, 1: {'image_id': 4, 'original_points3d': vrf_pts[:5]}}},
            allow_pickle=True)
    cfg = {'landmark_path': str(tmp_path), 'n_cluster': 8, 'cluster_mode': 'xz', 'cluster_method': 'birch', 'localization': {'threshold': 12}}
    smap = M.load_single_map(cfg, matcher=None, with_compress=True, device='cpu')
    assert smap.seg_ref_frame_ids == {3: [2, 4]}
    f2 = smap.reference_frames[2]
    assert np.array_equal(f2.point3D_ids, vrf_pts)   # the vrf dictionary overrides the frame's own point list
    desc = np.load(mdir / 'point3D_desc.npy', allow_pickle=True)[()]
    lab = dict(zip(labelled.tolist(), np.load(tmp_path / 'point3D_cluster_n8_xz_birch.npy', allow_pickle=True)[()]['label'].tolist()))
    assert np.allclose(f2.descriptors, np.array([desc[int(v)] for v in vrf_pts]))
    assert np.array_equal(f2.keypoint_segs, np.array([lab[int(v)] for v in vrf_pts]))
    # projection + score rule, against a direct restatement of refframe.py:99-147
    im, cam = imgs[2], cams[imgs[2].camera_id]
    K = np.eye(3); K[0, 0], K[1, 1], K[0, 2], K[1, 2] = cam.params[0], cam.params[1], cam.params[2], cam.params[3]
    T = np.eye(4); T[:3, :3] = M.qvec2rotmat(im.qvec); T[:3, 3] = im.tvec
    xyz = np.array([pts[int(v)].xyz for v in vrf_pts])
    uv = K @ (T @ np.hstack([xyz, np.ones((xyz.shape[0], 1))]).T)[:3]
    assert np.allclose(f2.keypoints[:, :2], (uv[:2] / uv[2]).T)
    assert np.allclose(f2.keypoints[:, 2], 1 / np.clip(np.array([float(pts[int(v)].error) for v in vrf_pts]) * 5, 1., 20.))
    for fid, fr in smap.reference_frames.items():     # only labelled points survive
        assert all(int(v) in lab for v in fr.point3D_ids) and fr.point3D_ids.size > 0
    d = f2.get_keypoints_by_sid(int(f2.keypoint_segs[0]))
    assert d['_device']['descriptors'].shape[1] == d['descriptors'].shape[0] == int((f2.keypoint_segs == f2.keypoint_segs[0]).sum())
