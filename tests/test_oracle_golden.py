"""Pin the CPU oracle (oracle/mvpy_oracle.py) against fixtures produced by the unmodified reference
(oracle/gen_golden.py).  CPU only; these are the known answers the CUDA path is later held to."""
import numpy as np
import pytest
import torch

from oracle import mvpy_oracle as orc

T = torch.from_numpy


def test_lag_window_and_delay_known_answer(golden):
    g = golden('delay.npz')
    assert np.array_equal(orc.lag_window(-0.03, 0.02, 100), g['window'])
    assert np.array_equal(orc.lag_window(0.01, 0.03, 100), g['window2'])          # abs(): -1..3
    assert np.array_equal(orc.lag_window(-1.0, 0.0, 200), g['window_cfg1'])
    Xt = orc.delay_design(T(g['X']), g['window'])
    assert torch.equal(Xt, T(g['Xt']))
    assert Xt[0].tolist() == [0, 0, 0, 0, 1, 2, 6, 6, 6, 6, 7, 8]               # clamped edges
    assert torch.equal(orc.delay_design(T(g['Xr']), g['window2']), T(g['Xt2']))
    assert torch.equal(orc.flatten_targets(T(g['yr'])), T(g['yt2']))


@pytest.mark.parametrize('solver', [orc.ridge_loo_fit, orc.ridge_loo_fit_primal, orc.ridge_loo_fit_chol])
def test_ridgecv_reference_cases(golden, solver):
    g = golden('ridgecv.npz')
    alphas = T(g['alphas'])
    for c in range(int(g['n_cases'])):
        n, p, f, icpt, per_target = g[f'c{c}_cfg']
        X, y = T(g[f'c{c}_X']), T(g[f'c{c}_y'])
        fit = solver(X, y, alphas, bool(icpt), bool(per_target))
        assert torch.equal(fit['alpha'].reshape(-1), T(g[f'c{c}_alpha']).reshape(-1)), c
        pred = orc.ridge_predict(X, fit['coef'], fit['intercept'])
        # the reference's own test pins predictions at rtol 1e-5 / atol 1e-7 (tests/estimators/test_ridgecv.py:87-92)
        np.testing.assert_allclose(pred.numpy(), g[f'c{c}_pred'], rtol=1e-6, atol=1e-8, err_msg=str(c))
        if n > p:  # coefficients are unique only for full column rank
            np.testing.assert_allclose(fit['coef'].numpy(), g[f'c{c}_coef'], rtol=1e-6, atol=1e-9, err_msg=str(c))


def test_ridgecv_fp32_tall(golden):
    g = golden('ridgecv.npz')
    X, y, al = T(g['f32_X']), T(g['f32_y']), T(g['f32_alphas'])
    for tag, pt in (('f32', False), ('f32pt', True)):
        fit = orc.ridge_loo_fit(X, y, al, True, pt)
        assert torch.equal(fit['alpha'].reshape(-1), T(g[tag + '_alpha']).reshape(-1))
        # the reference's fp32 SVD route itself sits ~1e-4 (abs) from the fp64 answer on this problem
        np.testing.assert_allclose(fit['coef'].numpy(), g[tag + '_coef'], rtol=1e-3, atol=3e-4)
        fit64 = orc.ridge_loo_fit_primal(X.double(), y.double(), al.double(), True, pt)
        np.testing.assert_allclose(fit64['coef'].numpy(), g[tag + '_coef'], rtol=1e-3, atol=3e-4)
        fit32p = orc.ridge_loo_fit_primal(X, y, al, True, pt)                       # kernels' route, fp32
        np.testing.assert_allclose(fit32p['coef'].numpy(), fit64['coef'].numpy(), rtol=1e-4, atol=1e-5)
        assert torch.equal(fit64['alpha'].float().reshape(-1), T(g[tag + '_alpha']).reshape(-1))


def test_scaler(golden):
    g = golden('scaler.npz')
    X = T(g['X'])
    for tag, dims in [('d0', (0,)), ('d01', (0, 1)), ('d2', (2,)), ('d02', (0, 2))]:
        st = orc.scaler_fit(X, dims)
        np.testing.assert_allclose(st['mean'].numpy(), g[tag + '_mean'], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(st['scale'].numpy(), g[tag + '_scale'], rtol=1e-6)
        np.testing.assert_allclose(orc.scaler_transform(X, st).numpy(), g[tag + '_tx'], rtol=1e-5, atol=1e-6)
    st = orc.scaler_fit(X, (0,))
    assert st['scale'][0, 2, 3] == 1.0                                            # zero variance -> 1
    np.testing.assert_allclose(orc.scaler_transform(T(g['Xnew']), st).numpy(), g['d0_tnew'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(orc.scaler_transform(X, st, with_std=False).numpy(), g['nostd_tx'], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(orc.scaler_transform(X, st, with_mean=False).numpy(), g['nomean_tx'], rtol=1e-6, atol=1e-6)


def test_kfold_bit_exact(golden):
    g = golden('kfold.npz')
    for n, k in [(10, 5), (11, 3), (120, 5), (7, 7)]:
        for i, (tr, te) in enumerate(orc.kfold_indices(n, k)):
            assert np.array_equal(tr.numpy(), g[f'n{n}k{k}_tr{i}']) and np.array_equal(te.numpy(), g[f'n{n}k{k}_te{i}'])
    for n, k, seed in [(20, 4, 0), (23, 5, 123)]:
        for i, (tr, te) in enumerate(orc.kfold_indices(n, k, shuffle=True, seed=seed)):
            assert np.array_equal(tr.numpy(), g[f'n{n}k{k}s{seed}_tr{i}'])
            assert np.array_equal(te.numpy(), g[f'n{n}k{k}s{seed}_te{i}'])
    with pytest.raises(ValueError):
        orc.kfold_indices(3, 5)


def test_metrics(golden):
    g = golden('metrics.npz')
    y, yh, yb = T(g['y']), T(g['yh']), T(g['yb'])
    np.testing.assert_allclose(orc.r2_last(y, yh).numpy(), g['r2'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(orc.r2_last(y, yh, yb).numpy(), g['r2_b'], rtol=1e-6, atol=1e-7)
    r = orc.pearsonr_last(y, yh).numpy()
    assert np.isnan(r[0, 0]) and np.isnan(g['pearsonr'][0, 0])
    np.testing.assert_allclose(r, g['pearsonr'], rtol=1e-5, atol=1e-6, equal_nan=True)
    assert orc.r2_last(y, yh)[0, 0] == 0.0


def test_trf_small_cross_val(golden):
    g = golden('trf_small.npz')
    X, y, al = T(g['X']), T(g['y']), T(g['alphas'])
    w = orc.lag_window(-0.4, 0.0, 50)
    out = orc.cross_val_trf(X, y, w, al, cv=5, metrics=('r2', 'pearsonr'))
    assert torch.equal(out['alpha'], T(g['cv_alpha']))
    np.testing.assert_allclose(out['r2'].numpy(), g['cv_score'], atol=2e-5)
    np.testing.assert_allclose(out['r2'].numpy(), g['cv2_r2'], atol=2e-5)
    np.testing.assert_allclose(out['pearsonr'].numpy(), g['cv2_pearsonr'], atol=2e-5)
    np.testing.assert_allclose(out['coef'].numpy(), g['cv_coef'], rtol=1e-3, atol=1e-6)
    # fp64 primal route (the kernels' algorithm) reproduces the reference's fp32 answers
    out64 = orc.cross_val_trf(X.double(), y.double(), w, al.double(), cv=5, solver=orc.ridge_loo_fit_primal)
    assert torch.equal(out64['alpha'].float(), T(g['cv_alpha']))
    np.testing.assert_allclose(out64['r2'].numpy(), g['cv_score'], atol=2e-5)
    # the Cholesky-per-alpha restatement (the fp64 checker of the BASELINE-size GPU tests) is the same function
    outc = orc.cross_val_trf(X.double(), y.double(), w, al.double(), cv=5, solver=orc.ridge_loo_fit_chol)
    assert torch.equal(outc['alpha'], out64['alpha'])
    np.testing.assert_allclose(outc['losses'].numpy(), out64['losses'].numpy(), rtol=1e-9)
    np.testing.assert_allclose(outc['coef'].numpy(), out64['coef'].numpy(), rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(outc['r2'].numpy(), g['cv_score'], atol=2e-5)


def test_trf_small_single_fit_and_patterns(golden):
    g = golden('trf_small.npz')
    X, y, al = T(g['X']), T(g['y']), T(g['alphas'])
    fit = orc.trf_fit(X[:32], y[:32], orc.lag_window(-0.4, 0.0, 50), al, patterns=True)
    assert torch.equal(fit['alpha'], T(g['fit_alpha']))
    np.testing.assert_allclose(fit['coef'].numpy(), g['fit_coef'], rtol=1e-3, atol=1e-6)
    # cov(y_t) of this 8-channel fixture has condition number ~6e14, so inv() amplifies fp32 rounding:
    # only a loose agreement is meaningful here (well-conditioned patterns are pinned in decoder.npz)
    np.testing.assert_allclose(fit['pattern'].numpy(), g['fit_pattern'], rtol=5e-2, atol=5e-2)
    pred = orc.trf_predict(X[32:], orc.lag_window(-0.4, 0.0, 50), fit['flat_coef'], fit['intercept'])
    np.testing.assert_allclose(pred.numpy(), g['fit_pred'], atol=2e-5)
    w2 = orc.lag_window(-0.1, 0.06, 50)
    fit2 = orc.trf_fit(X[:32], y[:32], w2, al, fit_intercept=False, alpha_per_target=True)
    assert torch.equal(fit2['alpha'], T(g['fit2_alpha']))
    np.testing.assert_allclose(orc.trf_predict(X[32:], w2, fit2['flat_coef'], fit2['intercept']).numpy(), g['fit2_pred'], atol=2e-5)


def test_hierarchical_and_shapley(golden):
    g = golden('trf_small.npz')
    X, y, al = T(g['X']), T(g['y']), T(g['alphas'])
    w = orc.lag_window(-0.4, 0.0, 50)
    sets, masks = orc.group_masks(torch.eye(4))
    assert np.array_equal(masks.numpy(), g['hier_mask'])
    assert [list(s) + [-1] * (4 - len(s)) for s in sets] == g['hier_sets'].tolist()
    groups = T(g['groups'])
    sets2, masks2 = orc.group_masks(groups)
    assert np.array_equal(masks2.numpy(), g['hier2_mask'])
    assert [list(s) + [-1] * (3 - len(s)) for s in sets2] == g['hier2_sets'].tolist()
    h = orc.hierarchical_trf(X, y, groups, w, al, cv=5)
    np.testing.assert_allclose(h['score'].numpy(), g['hier2_score'], atol=2e-5)
    torch.manual_seed(1)
    s = orc.shapley_trf(X, y, groups, w, al, n_permutations=3, cv=5)
    assert torch.equal(s['order'], T(g['shap_order']))                            # bit-exact permutation orders
    np.testing.assert_allclose(s['score'].numpy(), g['shap_score'], atol=4e-5)


def test_decoder_and_sliding(golden):
    g = golden('decoder.npz')
    X, y, al = T(g['X']), T(g['y']), T(g['alphas'])
    fits = orc.sliding_decoder_fit(X, y, al)
    for t, f in enumerate(fits):
        assert torch.equal(f['alpha'], T(g['alpha'][t]))
        np.testing.assert_allclose(f['coef'].numpy(), g['coef'][t], rtol=1e-3, atol=1e-5)
        np.testing.assert_allclose(f['coef'].numpy(), g['sliding_coef'][t], rtol=1e-3, atol=1e-5)
        np.testing.assert_allclose(f['pattern'].numpy(), g['pattern'][t], rtol=2e-3, atol=1e-4)
        np.testing.assert_allclose(orc.ridge_predict(X[..., t], f['coef'], f['intercept']).numpy(), g['pred'][..., t], rtol=1e-4, atol=3e-4)
    f1 = orc.decoder_fit(X[..., 0], y[:, 0, 0], al)
    assert torch.equal(f1['alpha'], T(g['single_alpha']))
    np.testing.assert_allclose(f1['coef'].numpy(), g['single_coef'], rtol=1e-3, atol=1e-5)


def test_encoder(golden):
    g = golden('encoder.npz')
    X, y, al = T(g['X']), T(g['y']), T(g['alphas'])
    f3 = orc.encoder_fit(X, y, al)
    assert torch.equal(f3['alpha'], T(g['alpha3']))
    np.testing.assert_allclose(f3['coef'].numpy(), g['coef3'], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(orc.encoder_predict(X, f3).numpy(), g['pred3'], atol=1e-4)
    f2 = orc.encoder_fit(X[..., 0], y[..., 0], al)
    assert torch.equal(f2['alpha'], T(g['alpha2']))
    np.testing.assert_allclose(orc.encoder_predict(X[..., 0], f2).numpy(), g['pred2'], atol=1e-4)


def test_rank_and_spearman(golden):
    g = golden('rank.npz')
    assert np.array_equal(orc.rank_last(T(g['x'])).numpy(), g['rank_x'])
    assert np.array_equal(orc.rank_last(T(g['xi'])).numpy(), g['rank_xi'])
    assert np.array_equal(orc.rank_last(torch.tensor([2, 0.5, 1, 1])).numpy(), g['doc'])
    np.testing.assert_allclose(orc.spearmanr_last(T(g['x']), T(g['y'])).numpy(), g['spearman'], atol=1e-6)


def test_receptive_field_reference_cases(golden):
    """the ReceptiveField restatement (direct lag sums) against the reference's FFT route -- per-epoch
    correlation blocks (with the reference's asymmetric lower blocks and head correction), penalties, alpha
    selection by k-fold / LOO, coefficients, intercepts, patterns, predictions."""
    from tests.rf_cases import RF_CASES
    g = golden('rf.npz')
    for name, ((t0, t1, fs), kw) in RF_CASES.items():
        X, y = T(g[f'{name}_X']), T(g[f'{name}_y'])
        X3, y3 = (X[:, None], y[:, None]) if X.ndim == 2 else (X, y)
        kw = dict(kw)
        alpha = kw.pop('alpha')
        if isinstance(alpha, list):
            alpha = torch.tensor(alpha)
        fit = orc.rf_fit(X3, y3, t0, t1, fs, alpha, **kw)
        np.testing.assert_allclose(fit['cov'].numpy(), g[f'{name}_cov'], rtol=1e-9, atol=1e-8, err_msg=name)
        assert float(fit['alpha']) == float(g[f'{name}_alpha']), name
        np.testing.assert_allclose(fit['coef'].numpy(), g[f'{name}_coef'], rtol=1e-7, atol=1e-9, err_msg=name)
        if kw.get('fit_intercept', True):
            np.testing.assert_allclose(fit['intercept'].numpy(), g[f'{name}_intercept'], rtol=1e-7, atol=1e-9, err_msg=name)
        else:
            assert float(g[f'{name}_intercept']) == 0.0
        if kw.get('patterns'):
            np.testing.assert_allclose(fit['pattern'].numpy(), g[f'{name}_pattern'], rtol=1e-7, atol=1e-9, err_msg=name)
        pred = orc.rf_predict(X3, fit['coef'], fit['intercept'], fit['s_min'])
        np.testing.assert_allclose(pred.numpy(), g[f'{name}_pred'], rtol=1e-4, atol=1e-4, err_msg=name)   # reference returns fp32
    from tests.rf_cases import RF_REG_CASES
    for i, (C, W, rt) in enumerate(RF_REG_CASES):
        assert np.array_equal(orc.rf_regulariser(C, W, rt).numpy(), g[f'reg{i}']), (C, W, rt)
    with pytest.raises(ValueError):
        orc.rf_regulariser(2, 3, 'lasso')


def test_pairwise_coupling(golden):
    """Wu-Lin coupling restatement against the reference's routine (classifier.py:573-653) incl. its last-write-wins updates."""
    g = golden('coupling.npz')
    for tag, K in (('a', 3), ('b', 5), ('c', 2), ('d', 4), ('e', 7)):
        p_pair = T(g[f'{tag}_pair'])
        eps = 5e-3 / K
        p = orc.pairwise_coupling(p_pair, max(100, K), eps, eps)
        tol = 1e-12 if p_pair.dtype == torch.float64 else 2e-6
        np.testing.assert_allclose(p.numpy(), g[f'{tag}_p'], rtol=0, atol=tol, err_msg=tag)
        np.testing.assert_allclose(p.sum(1).numpy(), 1.0, atol=1e-6)
