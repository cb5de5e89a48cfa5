"""ctypes binding of ``csrc/libpram_b200.so`` (the C ABI declared in ``include/pram_b200.h``).

There is no CPU fallback: if the shared library is missing or a call fails, this raises.
PyTorch is used only to own device memory and streams; every argument crossing the boundary is a
raw pointer / size / ``cudaStream_t``.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_HERE = Path(__file__).resolve().parent
# PRAM_LIB: another build of the same library (A/B timing of a compile-time variant); default = the in-tree build
LIB_PATH = Path(os.environ['PRAM_LIB']) if os.environ.get('PRAM_LIB') else _HERE / 'csrc' / 'libpram_b200.so'

_P = C.c_void_p
_I = C.c_int
_L = C.c_longlong
_F = C.c_float

# name -> (restype, argtypes); mirrors include/pram_b200.h one to one
SIGNATURES = {
    'pram_version': (_I, []),
    'pram_launch_count': (C.c_ulonglong, []),
    'pram_error_string': (C.c_char_p, [_I]),
    'pram_last_cuda_error': (C.c_char_p, []),
    'pram_set_launch_predicate': (_I, [_P]),
    'pram_set_pdl': (_I, [_I]),
    'pram_get_pdl': (_I, []),
    'pram_score_map': (_I, [_P, _L, _L, _L, _L, _I, _I, _I, _P, _P]),
    'pram_resize_bilinear': (_I, [_P, _I, _I, _I, _P, _I, _I, _P]),
    'pram_nms_candidates': (_I, [_P, _I, _I, _I, _I, _F, _F, _P, _P, _I, _P, _P, _P]),
    'pram_select_keypoints': (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _F, _F, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P]),
    'pram_sample_features': (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _P, _P]),
    'pram_gather_scores': (_I, [_P, _I, _I, _I, _P, _P, _I, _P, _P]),
    'pram_posenc': (_I, [_P, _I, _F, _F, _I, _P, _P, _P, _P]),
    'pram_conv_f32': (_I, [_P, _L, _P, _P, _P, _L, _P, _L, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'pram_linear_f32': (_I, [_P, _L, _P, _P, _P, _L, _P, _L, _L, _I, _I, _I, _I, _L, _L, _L, _P]),
    'pram_gconv3x3_f32': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'pram_l2norm_rows': (_I, [_P, _P, _L, _I, _P]),
    'pram_layernorm_gelu': (_I, [_P, _P, _P, _P, _L, _I, _I, _P]),
    'pram_rotary_split': (_I, [_P, _I, _I, _I, _I, _P, _P, _F, _P, _P, _P, _P]),
    'pram_attention_f32': (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _P, _I, _P, _P, _P]),
    'pram_attention_f32_colmean_ws_floats': (_L, [_I, _I, _I, _I]),
    'pram_sinkhorn_workspace_floats': (_L, [_I, _I, _I]),
    'pram_sinkhorn_match': (_I, [_P, _I, _I, _I, _P, _I, _F, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P]),
}



class TcArgs(C.Structure):
    """Mirror of ``struct pram_tc_args`` (include/pram_b200.h)."""
    _fields_ = [
        ('a_hi', _P), ('a_lo', _P), ('a_ld', _L),
        ('in_W', _I), ('in_H', _I), ('in_planes', _I), ('Cin', _I),
        ('w_hi', _P), ('w_lo', _P), ('w_planes', _I),
        ('B', _I), ('Ho', _I), ('Wo', _I), ('N', _I),
        ('tw_log2', _I), ('ntaps', _I),
        ('tap_dx', _I * 9), ('tap_dy', _I * 9), ('tap_plane', _I * 9),
        ('planes_per_image', _I), ('w_batch_mult', _I),
        ('bias', _P), ('res', _P), ('res_ld', _L), ('relu', _I),
        ('out_f32', _P), ('ld_f32', _L),
        ('out_hi', _P), ('out_lo', _P), ('ld_bf', _L),
        ('ps_hi', _P), ('ps_lo', _P), ('ld_ps', _L),
        ('l2norm', _I), ('split', _I), ('bn', _I),
        ('qkv_mode', _I), ('cosb', _P), ('sinb', _P), ('qk_scale', _F),
        ('q_hi', _P), ('q_lo', _P), ('k_hi', _P), ('k_lo', _P), ('v_hi', _P), ('v_lo', _P),
        ('seg_split', _I), ('seg_n0', _I), ('seg_n1', _I), ('heads', _I),
        ('cluster', _I), ('l2_prefetch', _I), ('f16', _I), ('res_hi', _P), ('res_lo', _P), ('v_f16', _I),
        ('out_h16', _P),
    ]


SIGNATURES['pram_gemm_tc'] = (_I, [C.POINTER(TcArgs), _P])


class MlpBlockArgs(C.Structure):
    """Mirror of ``struct pram_mlp_block_args`` (include/pram_b200.h)."""
    _fields_ = [
        ('a_hi', _P), ('a_lo', _P), ('lda', _L), ('T', _I),
        ('w1_hi', _P), ('w1_lo', _P), ('w3_hi', _P), ('w3_lo', _P), ('tables_host', _P),
        ('res', _P), ('res_ld', _L),
        ('out_f32', _P), ('ld_f32', _L),
        ('out_hi', _P), ('out_lo', _P), ('ld_bf', _L),
        ('split', _I), ('dbg', _P),
    ]


SIGNATURES['pram_mlp_block_tc'] = (_I, [C.POINTER(MlpBlockArgs), _P])
SIGNATURES['pram_split_f16'] = (_I, [_P, _P, _P, _L, _P])
SIGNATURES['pram_split_bf16'] = (_I, [_P, _P, _P, _L, _P])
SIGNATURES['pram_cast_f16'] = (_I, [_P, _P, _L, _P])
SIGNATURES['pram_attention_tc'] = (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _P, _P, _I, _I, _I, _I, _P, _P])
SIGNATURES['pram_attention_tc_shift'] = (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P])
SIGNATURES['pram_attention_tc_lse'] = (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _P, _P, _I, _I, _I, _I, _P, _P, _I, _P])
SIGNATURES['pram_attention_colsum_tc'] = (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _P, _I, _P, _I, _I, _I, _P, _P])
SIGNATURES['pram_colmean_reduce'] = (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P])
SIGNATURES['pram_adagml_prune'] = (_I, [_P, _I, _I, _I, _P, _P, _F, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P])
SIGNATURES['pram_adagml_move'] = (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _L, _P, _P, _L, _P, _P, _P, _P, _P, _P, _P])
SIGNATURES['pram_adagml_latch'] = (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P])
SIGNATURES['pram_adagml_scatter'] = (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P])
SIGNATURES['pram_attention_prep'] = (_I, [_P, _I, _I, _I, _I, _P, _P, _F, _P, _P, _P, _P, _P, _P, _I, _P])
_D = C.c_double
SIGNATURES['pram_ransac_workspace_bytes'] = (_L, [_I, _I, _I])
SIGNATURES['pram_ransac_pnp'] = (_I, [_P, _P, _P, _I, _I, _I, _D, _D, _D, _D, _D, _D, _I, _I, _I, _I, C.c_uint, _P, _P, _P, _P,
                                      _P, _P, _P])
SIGNATURES['pram_ransac_pnp_corr'] = (_I, [_P, _P, _I, _I, _D, _D, _I, _I, _I, _I, C.c_uint, _P, _P, _P, _P, _P, _P, _P])
SIGNATURES['pram_segmentation'] = (_I, [_P, _I, _I, _F, _P, _P, _P, _P, _P])
SIGNATURES['pram_rank_landmarks'] = (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P])
SIGNATURES['pram_project_points'] = (_I, [_P, _I, _P, _D, _D, _D, _D, _D, _D, _P, _P, _P])
SIGNATURES['pram_projection_top2'] = (_I, [_P, _I, _I, _I, _P, _P, _P, _F, _F, _P, _P, _P, _P])
SIGNATURES['pram_nn_match'] = (_I, [_P, _P, _I, _I, _I, _F, _F, _I, _P, _P, _P, _P, _P])
SIGNATURES['pram_gconv3x3_split'] = (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P])
SIGNATURES['pram_gconv3x3_tc'] = (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P])
SIGNATURES['pram_conv1a_tc'] = (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _I, _P])
SIGNATURES['pram_conv1a'] = (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P])
SIGNATURES['pram_layernorm_gelu_split'] = (_I, [_P, _P, _P, _P, _P, _P, _L, _I, _I, _P])

_lib = None


class PramError(RuntimeError):
    pass


def load(path: os.PathLike | None = None) -> C.CDLL:
    """Load the shared library (once) and attach argument types.  Fails loudly."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path is not None else LIB_PATH
    if not p.exists():
        raise PramError(f'{p} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                        f'(nvcc, sm_100a). There is no CPU fallback.')
    lib = C.CDLL(str(p))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def ptr(t: torch.Tensor | None):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def call(name: str, *args):
    """Invoke an int-returning ABI function and raise on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.pram_error_string(rc).decode()
        cuda = lib.pram_last_cuda_error().decode() if rc == -2 else ''
        raise PramError(f'{name} failed: {msg} {cuda}')
    return rc


def launch_count() -> int:
    return int(load().pram_launch_count())


def require_cuda(t: torch.Tensor, what: str = 'tensor'):
    if not t.is_cuda:
        raise PramError(f'{what} must live on a CUDA device: pram_b200 has no CPU path '
                        f'(the CPU oracle under oracle/ is test infrastructure only)')
