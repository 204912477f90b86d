"""The C-ABI library loads and exports every symbol include/mvb.h declares (no compute calls: CPU box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, 'include', 'mvb.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(mvb_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_expected_entry_points():
    syms = _header_symbols()
    for must in ['mvb_scaler_fit', 'mvb_trf_pad', 'mvb_trf_stats', 'mvb_eigh', 'mvb_ridge_solve', 'mvb_trf_predict',
                 'mvb_score', 'mvb_contract', 'mvb_last_error']:
        assert must in syms


def test_library_exports_every_declared_symbol():
    from mvpy_b200 import _lib, build
    build.build(force=False, verbose=False)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _header_symbols():
        assert hasattr(lib, name), f'{name} declared in mvb.h but not exported by libmvb.so'
    assert sorted(_lib._SIGNATURES) == _header_symbols()          # the ctypes table binds exactly the header
    assert _lib.load().mvb_version() >= 100


def test_argument_errors_are_reported_without_a_gpu():
    from mvpy_b200 import _lib
    lib = _lib.load()
    assert lib.mvb_scaler_fit(None, 4, 4, 0, None, None, None, None) == -1
    assert b'scaler_fit' in lib.mvb_last_error()
    assert lib.mvb_score(None, None, None, 1, 1, 0, None, None, None) == -1
    assert lib.mvb_trf_stats_workspace_bytes(None, 3) == 0


def test_no_cpu_fallback():
    import torch
    from mvpy_b200 import _lib
    import mvpy_b200 as mv
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_lib.MvbError):
        mv.estimators.RidgeCV(alphas=[1.0]).fit(torch.randn(10, 3), torch.randn(10, 2))
    with pytest.raises(_lib.MvbError):
        mv.math.r2(torch.randn(4, 5), torch.randn(4, 5))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'mvpy_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+[\w.]*oracle', src, flags=re.M), f'{f} imports the oracle'


def test_stats_workspace_carries_design_planes_for_large_fp32_problems():
    """mvb_trf_stats_workspace_bytes is host-only arithmetic: for fp32 designs with p >= 1024 columns it includes the
    pre-split TF32 planes of the design (8 bytes per entry of the 256-row x 32-k tiled lag expansion, the B operand of
    the Gram on the TMEM-A contraction kernel); fp64 and small designs do not pay for them."""
    from mvpy_b200 import _lib
    lib = _lib.load()

    def ws(n_trials, n_feat, n_time, n_lag, dtype, F=7):
        d = _lib.mvb_design(0x1000, n_feat * (n_time + n_lag - 1), n_time + n_lag - 1, n_time, n_lag, None, n_trials, None,
                            n_feat, dtype)          # xp is never dereferenced by the size query
        return lib.mvb_trf_stats_workspace_bytes(ctypes.byref(d), F)

    n, p = 12 * 400, 6 * 201
    tiles, stages = -(-p // 256), -(-n // 32)
    planes = tiles * stages * 2 * 8 * (256 * 16 + 16)
    big32, big64 = ws(12, 6, 400, 201, 0), ws(12, 6, 400, 201, 1)
    assert big32 >= planes and big32 - planes < 64 << 20
    assert big64 < planes                                    # fp64 runs on the CUDA-core contraction: no planes
    assert ws(12, 5, 400, 201, 0) < planes // 2              # p = 1005 < 1024: shared-memory kernel, no planes
    # BASELINE configs[1]: 51 tiles x 1200 stages x 65 792 B = 4.03 GB on top of the split-K scratch
    cfg2 = ws(96, 64, 400, 201, 0, F=306)
    assert 51 * 1200 * 65792 <= cfg2 < 51 * 1200 * 65792 + (2 << 30)
