"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a, loads, and exports every
symbol include/pram_b200.h declares (no compute calls without a GPU)."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / 'include' / 'pram_b200.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(pram_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_symbols():
    names = _declared()
    assert 'pram_sinkhorn_match' in names and 'pram_nms_candidates' in names
    assert len(names) >= 15


def test_library_exports_every_declared_symbol(lib):
    for name in _declared():
        assert hasattr(lib, name), f'{name} declared in include/pram_b200.h but not exported'


def test_binding_table_matches_header(lib):
    from pram_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    assert lib.pram_version() >= 100
    assert lib.pram_error_string(-1).decode() == 'invalid argument'


def test_missing_library_fails_loudly(tmp_path):
    import pytest
    from pram_b200 import _lib
    with pytest.raises(_lib.PramError):
        _lib.load(tmp_path / 'nope.so')


def test_no_cpu_fallback():
    """CPU tensors are rejected: the product path never routes through torch CPU or the oracle."""
    import pytest
    import torch
    from pram_b200 import _lib
    from pram_b200.nets.sfd2 import ResNet4x
    net = ResNet4x()
    with pytest.raises(_lib.PramError):
        net.extract_local_global({'image': torch.zeros(1, 3, 32, 32)})


def test_product_does_not_import_oracle():
    for p in (ROOT / 'pram_b200').rglob('*.py'):
        src = p.read_text()
        assert 'import oracle' not in src and 'from oracle' not in src, p
