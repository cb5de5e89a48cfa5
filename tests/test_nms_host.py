"""CPU check of the NMS tile kernel's logic: tests/host/nms_host.cu runs the SAME phase functions the CUDA kernel
runs (pram_b200/csrc/nms_tile.cuh), sequentially on the host, and the result must be bit-identical to the oracle's
simple_nms (reference nets/sfd2.py:20-35) -- including exact ties, plateaus, borders and ragged sizes."""
import ctypes
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import pram_oracle as O

HOST = Path(__file__).resolve().parent / 'host'


@pytest.fixture(scope='module')
def nms_host():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    so = HOST / '_nms_host.so'
    src = HOST / 'nms_host.cu'
    hdr = HOST.parents[1] / 'pram_b200' / 'csrc' / 'nms_tile.cuh'
    if not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run([nvcc, '-O2', '-std=c++17', '-Xcompiler', '-fPIC', '-shared', '-Wno-deprecated-gpu-targets', '-o', str(so), str(src)],
                       check=True, capture_output=True)
    lib = ctypes.CDLL(str(so))
    lib.nms_host.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.nms_host.restype = ctypes.c_int

    def run(score: torch.Tensor, radius: int, th: int) -> torch.Tensor:
        score = score.contiguous().float()
        out = torch.full_like(score, float('nan'))
        b, h, w = score.shape
        assert lib.nms_host(score.data_ptr(), b, h, w, radius, th, out.data_ptr()) == 0
        return out
    return run


def _maps(h, w, seed):
    g = torch.Generator().manual_seed(seed)
    smooth = torch.rand(1, h, w, generator=g)
    quant = torch.randint(0, 6, (1, h, w), generator=g).float() / 8      # many exact ties / plateaus
    sparse = torch.rand(1, h, w, generator=g) * (torch.rand(1, h, w, generator=g) > 0.97)  # exact zeros
    const = torch.full((1, h, w), 0.25)
    return torch.cat([smooth, quant, sparse, const], 0)


@pytest.mark.parametrize('h,w', [(480, 640), (97, 131), (24, 128), (7, 5), (200, 333)])
@pytest.mark.parametrize('radius', [4, 3])
@pytest.mark.parametrize('th', [96, 24])
def test_host_emulation_bit_exact(nms_host, h, w, radius, th):
    s = _maps(h, w, seed=h * 7 + w)
    ref = O.simple_nms(s, radius)
    got = nms_host(s, radius, th)
    assert torch.equal(got, ref), f'{(got != ref).sum().item()} pixels differ'


@pytest.mark.parametrize('radius', [0, 1, 2])
def test_host_emulation_small_radii(nms_host, radius):
    s = _maps(61, 150, seed=radius)
    assert torch.equal(nms_host(s, radius, 24), O.simple_nms(s, radius))


def test_host_emulation_golden(nms_host, golden):
    g = golden('sfd2_160x120.npz')
    score = torch.from_numpy(g['score_map'])
    for r, key in ((4, 'nms4'), (3, 'nms3')):
        for th in (96, 24):
            assert np.array_equal(nms_host(score, r, th).numpy(), g[key])
