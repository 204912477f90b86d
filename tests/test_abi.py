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


def test_stats_workspace_never_holds_the_lag_expansion(monkeypatch):
    """mvb_trf_stats_workspace_bytes is host-only arithmetic.  The Gram of a lag-expanded design is built from the padded
    stimulus itself: by the displacement-structure recurrence (gram_recurrence.cuh; scratch = the C first block rows) or, with
    MVB_GRAM_RECURRENCE=0, by the implicit-Toeplitz kernel (toeplitz_gram.cuh; scratch = the "pack": 2 TF32 planes x 4 shifted
    copies of the training series, 32 bytes per padded sample) -- never anything of the size of the lag expansion
    (n x p x 4 bytes = 1.98 GB at BASELINE configs[1]; timedelayed.py:415-423 materialises exactly that)."""
    from mvpy_b200 import _lib
    lib = _lib.load()

    def ws(n_trials, n_feat, n_time, n_lag, dtype, F=7):
        d = _lib.mvb_design(0x1000, n_feat * (n_time + n_lag - 1), n_time + n_lag - 1, n_time, n_lag, None, n_trials, None,
                            n_feat, dtype)          # xp is never dereferenced by the size query
        return lib.mvb_trf_stats_workspace_bytes(ctypes.byref(d), F)

    # BASELINE configs[1] (96 training trials x 64 features x 400 samples, 201 lags, 306 channels)
    expansion = 96 * 400 * 64 * 201 * 4
    monkeypatch.setenv('MVB_GRAM_RECURRENCE', '1')
    rec = ws(96, 64, 400, 201, 0, F=306)
    assert 64 * 12864 * 4 <= rec < (256 << 20) and rec < expansion // 4, (rec, expansion)
    monkeypatch.setenv('MVB_GRAM_RECURRENCE', '0')
    cfg2 = ws(96, 64, 400, 201, 0, F=306)
    pitch = -(-(400 + 200 + 64) // 4) * 4
    pack = 96 * 64 * pitch * 8 * 4
    assert pack <= cfg2 < (1 << 30) and cfg2 < expansion // 2, (cfg2, pack, expansion)
    assert ws(96, 64, 400, 201, 1, F=306) < cfg2 - pack + (64 << 20)     # fp64: CUDA-core contraction, no pack
    assert ws(96, 64, 404, 201, 0, F=306) < cfg2 - pack + (64 << 20)     # 404 samples: no whole MMA K steps per trial -> gather kernel, no pack
    assert ws(12, 5, 400, 201, 0) < (64 << 20)                           # small design: gather kernel with split-K
