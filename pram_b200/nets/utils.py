"""Drop-in for reference ``nets/utils.py`` (host-side helpers only)."""
import torch

eps = 1e-8


def arange_like(x, dim: int):
    return torch.arange(x.shape[dim], device=x.device, dtype=x.dtype)


def image_wh(image_shape):
    """(width, height) exactly as the reference unpacks it: ``_, _, height, width = image_shape``
    (nets/utils.py:19).  Callers that pass (1,3,W,H) (localization/singlemap3d.py:147) therefore get
    the swapped centre, which is reproduced, not fixed."""
    _, _, height, width = image_shape
    return float(width), float(height)


def normalize_keypoints(kpts: torch.Tensor, image_shape):
    """Reference nets/utils.py:17-24, on the device through the positional-encoding kernel's own
    arithmetic is not needed here: this helper is only for callers that want the normalised
    coordinates themselves (tiny elementwise torch plumbing)."""
    width, height = image_wh(image_shape)
    size = torch.tensor([width, height], device=kpts.device, dtype=kpts.dtype)
    return (kpts - size / 2) / (size.max() * 0.7)
