"""CPU tests of host-side logic that mirrors reference code paths outside the kernels."""
import numpy as np
import pytest

from oracle import pram_oracle as O


@pytest.mark.parametrize('topK', [-1, 5, 40, 70, 1000])
def test_select_with_mask_matches_reference_loops(topK):
    """Vectorised mask-guided keypoint selection (export path) vs the literal loop restatement of
    reference nets/sfd2.py:502-571, for every topK regime (fewer / more than the labelled ones, everything, off)."""
    from pram_b200.nets.sfd2 import select_with_mask
    rs = np.random.RandomState(topK + 7)
    h, w, n = 60, 80, 120
    mask = np.zeros((h, w, 3), np.uint8)
    mask[10:30, 5:40] = (3, 0, 0)
    mask[35:55, 30:70] = (1, 2, 0)   # id = 1 + 512
    kp = np.stack([rs.uniform(0, w - 1e-3, n), rs.uniform(0, h - 1e-3, n)], 1)
    sc = rs.rand(n)
    de = rs.randn(n, 8)
    ref = O.select_with_mask_loops(kp.copy(), sc.copy(), de.copy(), mask, topK)
    out = select_with_mask(kp, sc, de, mask, topK)
    assert set(out) == set(ref)
    for k in ref:
        assert out[k].dtype == ref[k].dtype and np.array_equal(out[k], ref[k]), k
    assert 0 < (out['labels'] != 0).sum() < n
