"""GPU PnP/RANSAC (K19) against synthetic known poses, the CPU restatement (oracle) and OpenCV.
Parity with pycolmap is unpinned: RANSAC is randomised, so the comparison is pose within
tolerance + the inlier mask recomputed from the returned pose."""
import numpy as np
import pytest
import torch

from oracle import pram_oracle as O

pytestmark = pytest.mark.gpu


def _scene(seed, n=300, outliers=0.2, noise=0.5, f=525.0, w=640, h=480):
    rs = np.random.RandomState(seed)
    ax = rs.randn(3); ax /= np.linalg.norm(ax)
    ang = np.deg2rad(rs.uniform(0, 15))
    q = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * ax])
    R = O.quat_to_rotmat(q)
    t = rs.uniform(-0.5, 0.5, 3)
    uv = np.stack([rs.uniform(4, w - 4, n), rs.uniform(4, h - 4, n)], 1)
    z = rs.uniform(1, 5, n)
    Xc = np.stack([(uv[:, 0] - w / 2) / f * z, (uv[:, 1] - h / 2) / f * z, z], 1)
    X = (Xc - t) @ R
    uv_n = uv + rs.normal(0, noise, uv.shape)
    out = rs.rand(n) < outliers
    uv_n[out] = np.stack([rs.uniform(0, w, out.sum()), rs.uniform(0, h, out.sum())], 1)
    cam = {'model': 'SIMPLE_PINHOLE', 'width': w, 'height': h, 'params': [f, w / 2, h / 2]}
    return uv_n, X, cam, q, t, ~out


@pytest.mark.parametrize('seed', range(5))
def test_known_pose(lib, dev, seed):
    from pram_b200.localization.pose_estimator import absolute_pose_estimation
    uv, X, cam, q, t, inl = _scene(seed)
    ret = absolute_pose_estimation(uv, X, cam, estimation_options={'ransac': {'max_error': 8.0}}, refinement_options={})
    assert ret is not None
    qv = ret['cam_from_world'].rotation.quat[[3, 0, 1, 2]]  # xyzw -> wxyz, as the reference does
    e_r, e_t = O.pose_error(qv, ret['cam_from_world'].translation, q, t)
    assert e_r < 0.3 and e_t < 0.03, (e_r, e_t)
    # inlier mask == mask recomputed from the returned pose (float64, same threshold rule)
    R = O.quat_to_rotmat(qv)
    f, cx, cy = cam['params']
    x = np.stack([(uv[:, 0] - cx) / f, (uv[:, 1] - cy) / f], 1)
    e = O._reproj_sq_err(R, ret['cam_from_world'].translation, x, X)
    assert np.array_equal(ret['inliers'], e <= (8.0 / f) ** 2)
    assert ret['num_inliers'] == ret['inliers'].sum()
    assert (ret['inliers'] == inl).mean() > 0.97
    # agreement with the CPU restatement and with OpenCV (independent implementation)
    ref = O.absolute_pose_estimation(uv, X, cam, max_error=8.0, max_num_trials=1000)
    e_r2, e_t2 = O.pose_error(qv, ret['cam_from_world'].translation, ref['qvec'], ref['tvec'])
    assert e_r2 < 0.3 and e_t2 < 0.03
    import cv2
    K = np.array([[f, 0, cx], [0, f, cy], [0, 0, 1.0]])
    ok, rvec, tvec, _ = cv2.solvePnPRansac(X, uv, K, None, reprojectionError=8.0, iterationsCount=1000, flags=cv2.SOLVEPNP_P3P)
    assert ok
    Rcv = cv2.Rodrigues(rvec)[0]
    e_r3, e_t3 = O.pose_error(qv, ret['cam_from_world'].translation, O.rotmat_to_quat(Rcv), tvec.ravel())
    assert e_r3 < 1.0 and e_t3 < 0.1


def test_failure_conventions(lib, dev):
    from pram_b200.localization.pose_estimator import absolute_pose_estimation
    cam = {'model': 'PINHOLE', 'width': 640, 'height': 480, 'params': [500, 500, 320, 240]}
    assert absolute_pose_estimation(np.zeros((2, 2)), np.zeros((2, 3)), cam) is None  # < 3 points -> None
    rs = np.random.RandomState(0)  # pure noise: no consensus -> few inliers, never a crash
    ret = absolute_pose_estimation(rs.uniform(0, 640, (50, 2)), rs.randn(50, 3), cam, {'ransac': {'max_error': 1.0}}, {})
    assert ret is None or ret['num_inliers'] < 15


def test_batched_device_api(lib, dev):
    """Frame-batched entry used by the runner: matches index into a reference set, -1 = unmatched."""
    from pram_b200 import ops
    B, n = 4, 256
    kp, xyz, mt, gt = [], [], [], []
    for b in range(B):
        uv, X, cam, q, t, inl = _scene(10 + b, n=n)
        perm = np.random.RandomState(b).permutation(n)
        m = np.argsort(perm)  # keypoint i <-> reference m[i]
        m[::7] = -1
        kp.append(uv - 0.5); xyz.append(X[perm]); mt.append(m); gt.append((q, t))
    out = ops.ransac_pnp(torch.tensor(np.stack(kp), device=dev).float(), torch.tensor(np.stack(mt), device=dev),
                         torch.tensor(np.stack(xyz), device=dev).float(), 525.0, 525.0, 320.0, 240.0, 8.0, pixel_shift=0.5)
    assert out['success'].all()
    for b in range(B):
        e_r, e_t = O.pose_error(out['qvec'][b].cpu().numpy(), out['tvec'][b].cpu().numpy(), *gt[b])
        assert e_r < 0.3 and e_t < 0.03, (b, e_r, e_t)
        assert not out['inliers'][b].cpu().numpy()[::7].any()  # unmatched keypoints are never inliers
