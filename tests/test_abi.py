"""The C-ABI library builds, loads and exports every symbol include/egopose_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built():
    from egopose_b200 import build
    return build.build()


def _declared():
    src = open(os.path.join(ROOT, 'include', 'egopose_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(egp_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(built)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_binding_covers_header(built):
    from egopose_b200 import lib
    assert sorted(lib.SYMBOLS) == _declared()
    L = lib.load()
    assert L.egp_version() == 100
    assert L.egp_gae_work_bytes(1024 * 10) > 0


def test_sass_is_sm100a(built):
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', built], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_no_oracle_import_in_product():
    """the product path must never route through the CPU oracle"""
    pkg = os.path.join(ROOT, 'egopose_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, f)).read()
                assert 'oracle' not in txt.replace('no oracle', ''), os.path.join(dp, f)


def test_compute_fails_loudly_without_gpu():
    import torch
    from egopose_b200 import lib
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from egopose_b200.mjcf import load_builtin
    import helpers
    with pytest.raises(lib.EgpError):
        lib.Model.from_cfg_dict(load_builtin(), helpers.cfg_dict())
    with pytest.raises(lib.EgpError):
        lib.gae(torch.zeros(4, dtype=torch.float64), torch.zeros(4, dtype=torch.float64),
                torch.zeros(4, dtype=torch.float64), 0.95, 0.95)
