import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / 'tests' / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(GOLDEN / name, allow_pickle=False))
    return load


@pytest.fixture(scope='session')
def lib():
    """Build (if stale) and load the C-ABI library."""
    from pram_b200 import build, _lib
    build.build()
    return _lib.load()


@pytest.fixture(scope='session')
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


@pytest.fixture()
def ref_loc(monkeypatch):
    """The reference's OWN ``localization.*`` modules, importable here with stub modules for the third-party packages
    that are absent (pycolmap, h5py, ...), and with ``Tensor.cuda()`` turned into a no-op for the duration of the test
    (the reference hard-codes ``.cuda()`` at its matcher call sites).  Build container only: skips when /root/reference
    is not mounted.  ``ref_loc.pycolmap`` is the stub module the reference code calls into -- tests assign
    ``absolute_pose_estimation`` on it."""
    import types
    import torch
    from oracle import ref_loader as RL
    if not RL.reference_available():
        pytest.skip('reference tree not mounted')
    RL.import_reference()

    class _Any(types.ModuleType):  # stub module: any attribute resolves to a dummy class
        def __getattr__(self, item):
            if item.startswith('__'):
                raise AttributeError(item)
            return type(item, (), {})
    for name in ('pycolmap', 'h5py', 'progressbar', 'open3d', 'tensorboardX', 'pypangolin', 'OpenGL', 'OpenGL.GL'):
        sys.modules.setdefault(name, _Any(name))
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    try:
        import localization.singlemap3d as singlemap3d
        import localization.tracker as tracker
        import localization.pose_estimator as pose_estimator
        import localization.extract_features as extract_features
        import localization.match_features_batch as match_features_batch
        import localization.refframe as refframe
        import localization.frame as frame
    except Exception as e:  # noqa: BLE001 -- the reference pulls many optional dependencies
        pytest.skip(f'reference localization modules not importable here: {e!r}')
    return types.SimpleNamespace(singlemap3d=singlemap3d, tracker=tracker, pose_estimator=pose_estimator,
                                 extract_features=extract_features, match_features_batch=match_features_batch,
                                 refframe=refframe, frame=frame, pycolmap=sys.modules['pycolmap'])
