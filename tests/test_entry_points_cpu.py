"""CPU tests of the entry-point modules' host side: the h5-like stores, the extract_features configuration / dataset /
post-processing against the reference's module, camera models of the pose operator against the oracle."""
import numpy as np
import pytest

from oracle import pram_oracle as O


def test_array_store_round_trip(tmp_path):
    from pram_b200.localization.h5store import ArrayStore, open_store
    p = tmp_path / 'feats-sfd2.h5'
    rs = np.random.RandomState(0)
    recs = {f'seq/{i}.png': {'keypoints': rs.rand(5 + i, 2), 'descriptors': rs.randn(128, 5 + i), 'scores': rs.rand(5 + i),
                             'image_size': np.array([640, 480])} for i in range(3)}
    fd = open_store(p, 'a', backend='npz')
    assert isinstance(fd, ArrayStore)
    for name, rec in recs.items():
        grp = fd.create_group(name)
        for k, v in rec.items():
            grp.create_dataset(k, data=v)
    with pytest.raises(ValueError):
        fd.create_group('seq/0.png')
    fd.close()
    fd = open_store(p, 'r')            # backend recognised from the file
    assert isinstance(fd, ArrayStore) and set(fd.keys()) == set(recs) and 'seq/1.png' in fd and 'nope' not in fd
    for name, rec in recs.items():
        for k, v in rec.items():
            got = fd[name][k][()]
            assert got.dtype == v.dtype and np.array_equal(got, v)
            assert np.array_equal(fd[name][k].__array__(), v) and fd[name][k].shape == v.shape
        assert dict(fd[name].items()).keys() == rec.keys()
    with pytest.raises(OSError):
        fd.create_group('x')
    fd.close()
    # append + delete + dtypes of the match records (int16 / float16)
    fd = open_store(p, 'a')
    del fd['seq/0.png']
    g = fd.create_group('a-b.png/c-d.png')
    g.create_dataset('matches0', data=np.array([1, -1, 3], np.int16))
    g.create_dataset('matching_scores0', data=np.array([0.5, 0, 0.25], np.float16))
    fd.close()
    with open_store(p, 'r') as fd:
        assert 'seq/0.png' not in fd and fd['a-b.png/c-d.png']['matches0'][()].dtype == np.int16
        assert fd['a-b.png/c-d.png']['matching_scores0'][()].dtype == np.float16
        assert np.array_equal(fd['seq/2.png']['keypoints'][()], recs['seq/2.png']['keypoints'])
    with pytest.raises(FileNotFoundError):
        open_store(tmp_path / 'missing.h5', 'r')


def test_extract_features_config_dataset_and_export_vs_reference(ref_loc, tmp_path):
    import cv2
    from pram_b200.localization import extract_features as E
    R = ref_loc.extract_features
    for name in E.confs:
        ours, theirs = E.confs[name], R.confs[name]
        assert ours == theirs, name
    assert set(R.confs) - set(E.confs) == {'superpoint-n4096'}   # SuperPoint is outside the hot path
    with pytest.raises(ValueError):
        E.get_model('superpoint', 'x')
    # dataset: same names, sizes and pixel values as the reference loader, with and without resizing
    rs = np.random.RandomState(1)
    (tmp_path / 'a').mkdir()
    cv2.imwrite(str(tmp_path / 'a' / 'one.png'), (rs.rand(60, 80, 3) * 255).astype(np.uint8))
    cv2.imwrite(str(tmp_path / 'two.jpg'), (rs.rand(50, 90, 3) * 255).astype(np.uint8))
    for pre in ({'grayscale': False, 'resize_max': False}, {'grayscale': False, 'resize_max': 64}, {'grayscale': True, 'resize_max': None}):
        ds_o, ds_r = E.ImageDataset(tmp_path, pre), R.ImageDataset(tmp_path, pre)
        assert len(ds_o) == len(ds_r) == 2
        by_name = {ds_r[i]['name']: ds_r[i] for i in range(2)}
        for i in range(2):
            d = ds_o[i]
            r = by_name[d['name']]
            assert np.array_equal(d['original_size'], r['original_size']) and d['image'].dtype == r['image'].dtype
            assert np.array_equal(d['image'], r['image'])
    lst = tmp_path / 'list.txt'
    lst.write_text('a/one.png\n')
    assert [str(p) for p in E.ImageDataset(tmp_path, {}, image_list=str(lst)).paths] == ['a/one.png']
    # per-image post-processing == the statements of the reference's main loop (:226-236)
    pred = {'keypoints': rs.rand(7, 2) * 60, 'scores': rs.rand(7), 'descriptors': rs.randn(7, 128)}
    out = E.export_one(pred, (1, 3, 48, 64), np.array([80, 60]))
    size = np.array((1, 3, 48, 64)[-2:][::-1])
    scales = (np.array([80, 60]) / size).astype(np.float32)
    assert np.array_equal(out['keypoints'], (pred['keypoints'] + .5) * scales[None] - .5)
    assert out['descriptors'].shape == (128, 7) and np.array_equal(out['image_size'], [80, 60])


CAMS = [
    {'model': 'SIMPLE_PINHOLE', 'width': 640, 'height': 480, 'params': [500.0, 320.0, 240.0]},
    {'model': 'PINHOLE', 'width': 640, 'height': 480, 'params': [500.0, 510.0, 320.0, 240.0]},
    {'model': 'SIMPLE_RADIAL', 'width': 640, 'height': 480, 'params': [500.0, 320.0, 240.0, -0.12]},
    {'model': 'RADIAL', 'width': 640, 'height': 480, 'params': [500.0, 320.0, 240.0, -0.1, 0.03]},
    {'model': 'OPENCV', 'width': 640, 'height': 480, 'params': [500.0, 505.0, 320.0, 240.0, -0.1, 0.02, 1e-3, -2e-3]},
]


@pytest.mark.parametrize('cam', CAMS, ids=[c['model'] for c in CAMS])
def test_camera_models_vs_oracle(cam):
    """Pixels -> camera plane: our damped fixed-point inverse against the oracle's Newton iteration (COLMAP's method) and
    against the forward lens model (round trip), float64."""
    from types import SimpleNamespace
    from pram_b200.localization import pose_estimator as P
    rs = np.random.RandomState(0)
    uv = np.stack([rs.uniform(-0.6, 0.6, 500), rs.uniform(-0.45, 0.45, 500)], 1)
    px = O.img_from_cam(cam, uv)
    ours = P.cam_from_img(cam, px)
    assert np.abs(ours - uv).max() < 1e-10
    assert np.abs(ours - O.cam_from_img(cam, px)).max() < 1e-10
    assert P.camera_intrinsics(cam) == O.camera_intrinsics(cam)
    # object-style cameras (namedtuple / pycolmap-like with an enum model) resolve the same way
    obj = SimpleNamespace(model=SimpleNamespace(name=cam['model']), params=cam['params'], width=640, height=480)
    assert P.camera_intrinsics(obj) == P.camera_intrinsics(cam)


def test_unsupported_camera_models_raise():
    from pram_b200.localization import pose_estimator as P
    fov = {'model': 'FOV', 'width': 640, 'height': 480, 'params': [500.0, 501.0, 320.0, 240.0, 0.9]}
    assert P.camera_intrinsics(fov) == (500.0, 501.0, 320.0, 240.0)      # layout known: never (p0, p0, p1, p2)
    with pytest.raises(ValueError):
        P.cam_from_img(fov, np.zeros((1, 2)))                            # ... but its lens model is not silently ignored
    with pytest.raises(ValueError):
        P.camera_intrinsics({'model': 'MY_MODEL', 'params': [1, 2, 3]})
