"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python oracle/gen_golden.py
The GPU box has no /root/reference; tests there read the committed .npz files.  Everything is
seeded; inputs are stored next to the outputs so neither side has to regenerate them.

Fixtures (all produced through the reference's public classes/functions):
  ridgecv.npz      _RidgeCV_torch on the reference's own test sizes (tests/estimators/test_ridgecv.py:
                   10x20 "gram", 20x10 "svd"; 1/3 targets; +-intercept; per-target alphas), fp64,
                   plus fp32 cases with more rows.
  delay.npz        _delay_matrices known answer (arange(24).reshape(2,2,6), lags -3..2) + a random case.
  scaler.npz       _Scaler_torch statistics/transforms for several dims.
  kfold.npz        KFold index listings (plain + shuffled with seeds).
  splitters.npz    RepeatedKFold and LeaveGroupsOut (single- and multi-label groups) index listings.
  stratified.npz   StratifiedKFold / RepeatedStratifiedKFold index listings (multi-feature labels, 3-D labels, shuffled with
                   the global generator seeded as stored in the script).
  metrics.npz      math.r2 / math.pearsonr incl. zero-variance rows and the y_b baseline.
  rank.npz         math.rank (ties, integer input, the docstring example) and math.spearmanr.
  trf_small.npz    make_meeg_continuous(fs=50, n_trials=40, n_channels=8, n_features=4) seed 0:
                   cross_val_score / hierarchical_score / shapley_score through Scaler+TimeDelayed.
  decoder.npz      per-slice RidgeCV(alpha_per_target) + Haufe pattern (RidgeDecoder maths), Sliding.
  b2b.npz          back-to-back regression composed from _RidgeCV_torch + _Scaler_torch as b2b.py:267-306 does.
  rf.npz           ReceptiveField (needs MVPY_COMPILE_TORCH=0): single alpha, k-fold and LOO alpha selection, laplacian
                   penalties, patterns, no edge correction / intercept, causal-only and anticausal-only lag ranges.
  clf.npz          LabelBinariser, math.accuracy / roc_auc, the Roc_auc metric (reference), RidgeClassifier OvR / OvO composed
                   from the reference's binariser and _RidgeCV_torch.
  coupling.npz     Wu-Lin pairwise coupling (classifier.py:_pairwise_coupling_torch, needs MVPY_COMPILE_TORCH=0) on random pairwise
                   probabilities, and the one-versus-one predict_proba flow of classifier.py:913-1000 on the clf.npz problem.
  encoder.npz      RidgeEncoder 2-D and 3-D (expanded design).
  cfg1_summary.npz config 1 (README flow, fs=200): per-fold alphas, mean scores, strided score samples.
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings('ignore')
sys.path.insert(0, '/root/reference')
import mvpy as ref  # noqa: E402
from mvpy.estimators.ridgecv import _RidgeCV_torch  # noqa: E402
from mvpy.estimators.timedelayed import _TimeDelayed_torch  # noqa: E402
from sklearn.pipeline import make_pipeline  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(8)


def npy(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


def save(name, **arrays):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: npy(v) for k, v in arrays.items()})
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB, {len(arrays)} arrays')


# ---------------------------------------------------------------------------------------- ridgecv
def regression_problem(n_samples, n_features, n_targets, snr=10.0):
    """the reference's own fixture (tests/conftest.py:56-85), default_rng(42)."""
    rng = np.random.default_rng(42)
    X = rng.normal(size=(n_samples, n_features))
    B = rng.normal(size=(n_features, n_targets))
    y = X @ B
    y = y + rng.normal(size=y.shape) * (y.std() / snr)
    return X, y


def gen_ridgecv():
    out = {}
    alphas10 = np.logspace(-5, 5, 10)
    cid = 0
    for (n, p) in [(10, 20), (20, 10), (60, 7), (25, 25)]:
        for f in (1, 3):
            for icpt in (True, False):
                for per_target in (False, True):
                    X, y = regression_problem(n, p, f)
                    m = _RidgeCV_torch(alphas=torch.from_numpy(alphas10), fit_intercept=icpt,
                                       alpha_per_target=per_target).fit(torch.from_numpy(X), torch.from_numpy(y))
                    pred = m.predict(torch.from_numpy(X))
                    key = f'c{cid}'
                    out[key + '_X'], out[key + '_y'] = X, y
                    out[key + '_cfg'] = np.array([n, p, f, int(icpt), int(per_target)])
                    out[key + '_coef'], out[key + '_alpha'] = m.coef_, m.alpha_
                    out[key + '_intercept'] = m.intercept_ if icpt else np.zeros((1, f))
                    out[key + '_pred'] = pred
                    cid += 1
    out['alphas'] = alphas10
    out['n_cases'] = np.array(cid)
    # fp32, tall problems (the regime of the TRF path)
    g = torch.Generator().manual_seed(7)
    X = torch.randn(400, 24, generator=g)
    B = torch.randn(24, 5, generator=g)
    y = X @ B + 3.0 * torch.randn(400, 5, generator=g) + 0.7
    X = X + 0.3
    al = torch.logspace(-3, 4, 12)
    for per_target in (False, True):
        m = _RidgeCV_torch(alphas=al, alpha_per_target=per_target).fit(X, y)
        tag = 'f32pt' if per_target else 'f32'
        out[tag + '_coef'], out[tag + '_alpha'], out[tag + '_intercept'] = m.coef_, m.alpha_, m.intercept_
    out['f32_X'], out['f32_y'], out['f32_alphas'] = X, y, al
    save('ridgecv.npz', **out)


# ------------------------------------------------------------------------------------------ delay
def gen_delay():
    td = _TimeDelayed_torch(-0.03, 0.02, 100, alphas=torch.tensor([1.0]))
    X = torch.arange(24).reshape(2, 2, 6).float()
    Xt = td._delay_matrices(X)
    g = torch.Generator().manual_seed(3)
    Xr, yr = torch.randn(3, 2, 9, generator=g), torch.randn(3, 4, 9, generator=g)
    td2 = _TimeDelayed_torch(0.01, 0.03, 100, alphas=torch.tensor([1.0]))
    Xt2, yt2 = td2._delay_matrices(Xr, yr)
    save('delay.npz', X=X, Xt=Xt, window=td.window, Xr=Xr, yr=yr, Xt2=Xt2, yt2=yt2, window2=td2.window,
         window_cfg1=_TimeDelayed_torch(-1.0, 0.0, 200, alphas=torch.tensor([1.0])).window)


# ----------------------------------------------------------------------------------------- scaler
def gen_scaler():
    g = torch.Generator().manual_seed(11)
    X = torch.randn(12, 5, 7, generator=g) * 3 + 2
    X[:, 2, 3] = 4.0  # zero-variance cell -> scale forced to 1
    out = dict(X=X)
    Xn = torch.randn(4, 5, 7, generator=g)
    out['Xnew'] = Xn
    for tag, dims in [('d0', None), ('d01', (0, 1)), ('d2', (2,)), ('d02', (0, 2))]:
        sc = ref.preprocessing.Scaler(dims=dims).to_torch().fit(X)
        out[tag + '_mean'], out[tag + '_var'], out[tag + '_scale'] = sc.mean_, sc.var_, sc.scale_
        out[tag + '_tx'] = sc.transform(X)
        if tag == 'd0':
            out[tag + '_tnew'] = sc.transform(Xn)
            out[tag + '_inv'] = sc.inverse_transform(sc.transform(Xn))
    sc = ref.preprocessing.Scaler(with_std=False).to_torch().fit(X)
    out['nostd_tx'] = sc.transform(X)
    sc = ref.preprocessing.Scaler(with_mean=False).to_torch().fit(X)
    out['nomean_tx'] = sc.transform(X)
    save('scaler.npz', **out)


# ------------------------------------------------------------------------------------------ kfold
def gen_kfold():
    out = {}
    for n, k in [(10, 5), (11, 3), (120, 5), (7, 7)]:
        X = torch.zeros(n, 2)
        for i, (tr, te) in enumerate(ref.crossvalidation.KFold(n_splits=k).split(X)):
            out[f'n{n}k{k}_tr{i}'], out[f'n{n}k{k}_te{i}'] = tr, te
    for n, k, seed in [(20, 4, 0), (23, 5, 123)]:
        X = torch.zeros(n, 2)
        for i, (tr, te) in enumerate(ref.crossvalidation.KFold(n_splits=k, shuffle=True, random_state=seed).split(X)):
            out[f'n{n}k{k}s{seed}_tr{i}'], out[f'n{n}k{k}s{seed}_te{i}'] = tr, te
    save('kfold.npz', **out)


# ---------------------------------------------------------------------------------------- metrics
def gen_splitters():
    out = {}
    X = torch.zeros(11, 2)
    rkf = ref.crossvalidation.RepeatedKFold(n_splits=3, n_repeats=2, random_state=7)
    for i, (tr, te) in enumerate(rkf.split(X)):
        out[f'rkf_tr{i}'], out[f'rkf_te{i}'] = tr, te
    obj = ref.crossvalidation.RepeatedKFold(n_splits=2, n_repeats=2, random_state=1).to_torch()
    for rep in range(2):                          # a kept torch object continues its random stream across calls
        for i, (tr, te) in enumerate(obj.split(torch.zeros(6))):
            out[f'rkfobj{rep}_tr{i}'], out[f'rkfobj{rep}_te{i}'] = tr, te
    g1 = torch.arange(4).repeat(3)
    lgo = ref.crossvalidation.LeaveGroupsOut(n_splits=2, n_repeats=2, random_state=3)
    for i, (tr, te) in enumerate(lgo.split(torch.zeros(12, 2), groups=g1)):
        out[f'lgo1_tr{i}'], out[f'lgo1_te{i}'] = tr, te
    g2 = torch.stack([torch.arange(12) % 4, (torch.arange(12) // 3) % 4], dim=1)
    lgo2 = ref.crossvalidation.LeaveGroupsOut(n_splits=2, n_repeats=1, random_state=5)
    for i, (tr, te) in enumerate(lgo2.split(torch.zeros(12, 2), groups=g2)):
        out[f'lgo2_tr{i}'], out[f'lgo2_te{i}'] = tr, te
    out['g1'], out['g2'] = g1, g2
    save('splitters.npz', **out)


def gen_stratified():
    out = {}
    g = torch.Generator().manual_seed(11)
    y1 = torch.tensor([2, 0, 1, 1, 0, 2, 2, 1, 0, 0, 1, 2, 0, 1, 2, 2, 0])            # one feature, three classes, uneven
    y2 = torch.stack([torch.randint(0, 2, (40,), generator=g), 10 + torch.randint(0, 3, (40,), generator=g)], dim=1)
    y3 = y2[:, :, None].expand(-1, -1, 5).float()                                        # (n, features, time): time 0 is used
    out['y1'], out['y2'], out['y3'] = y1, y2, y3
    for name, y, k in (('s1', y1, 3), ('s2', y2, 4), ('s3', y3, 4)):
        skf = ref.crossvalidation.StratifiedKFold(n_splits=k)
        for i, (tr, te) in enumerate(skf.split(torch.zeros(y.shape[0], 2), y)):
            out[f'{name}_tr{i}'], out[f'{name}_te{i}'] = tr, te
    torch.manual_seed(21)                            # the reference's shuffle draws from torch's global generator
    skf = ref.crossvalidation.StratifiedKFold(n_splits=4, shuffle=True, random_state=3)
    for i, (tr, te) in enumerate(skf.split(torch.zeros(40, 2), y2)):
        out[f'sh_tr{i}'], out[f'sh_te{i}'] = tr, te
    torch.manual_seed(22)
    rsk = ref.crossvalidation.RepeatedStratifiedKFold(n_splits=3, n_repeats=2, random_state=5)
    for i, (tr, te) in enumerate(rsk.split(torch.zeros(17, 2), y1)):
        out[f'rsk_tr{i}'], out[f'rsk_te{i}'] = tr, te
    save('stratified.npz', **out)


def gen_rank():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(5, 7, 24, generator=g)
    x[0, 0, 3] = x[0, 0, 9] = x[0, 0, 17]                         # a three-way tie and a pair
    x[2, 4, 0] = x[2, 4, 23]
    xi = torch.randint(0, 6, (4, 30), generator=g)                # integer input with many ties
    y = 0.6 * x + 0.8 * torch.randn(5, 7, 24, generator=g)
    save('rank.npz', x=x, xi=xi, y=y, rank_x=ref.math.rank(x), rank_xi=ref.math.rank(xi),
         doc=ref.math.rank(torch.tensor([2, 0.5, 1, 1])), spearman=ref.math.spearmanr(x, y),
         spearman_d=ref.math.spearmanr_d(x, y), x64=x.double(), rank_x64=ref.math.rank(x.double()))


def gen_metrics():
    g = torch.Generator().manual_seed(5)
    y = torch.randn(6, 9, 24, generator=g)
    yh = y + 0.5 * torch.randn(6, 9, 24, generator=g)
    y[0, 0] = 1.5  # zero variance row: r2 -> 0 (guard), pearson -> nan
    yb = torch.randn(6, 9, 1, generator=g) * 0.1
    save('metrics.npz', y=y, yh=yh, yb=yb, r2=ref.math.r2(y, yh), r2_b=ref.math.r2(y, yh, yb),
         pearsonr=ref.math.pearsonr(y, yh))


# -------------------------------------------------------------------------------------- trf small
def gen_trf_small():
    torch.manual_seed(0)
    X, y = ref.dataset.make_meeg_continuous(fs=50, n_trials=40, n_channels=8, n_features=4)
    alphas = torch.logspace(-5, 5, 20)

    def pipe():
        return make_pipeline(ref.preprocessing.Scaler().to_torch(),
                             ref.estimators.TimeDelayed(-0.4, 0.0, 50, alphas=alphas))

    out = dict(X=X, y=y, alphas=alphas)
    val, sc = ref.crossvalidation.cross_val_score(pipe(), X, y, cv=5)
    out['cv_score'] = sc
    out['cv_alpha'] = torch.stack([m[-1].estimator.alpha_ for m in val.model_])
    out['cv_coef'] = torch.stack([m[-1].coef_ for m in val.model_])
    out['cv_intercept'] = torch.stack([m[-1].intercept_ for m in val.model_])
    out['cv_scaler_mean'] = torch.stack([m[0].mean_ for m in val.model_])
    out['cv_scaler_scale'] = torch.stack([m[0].scale_ for m in val.model_])
    # two metrics at once (dict output path)
    _, sc2 = ref.crossvalidation.cross_val_score(pipe(), X, y, cv=5, metric=(ref.metrics.r2, ref.metrics.pearsonr))
    out['cv2_r2'], out['cv2_pearsonr'] = sc2['r2'], sc2['pearsonr']
    # bare estimator without the scaler, single fit: coefficients, patterns, predictions
    td = ref.estimators.TimeDelayed(-0.4, 0.0, 50, alphas=alphas, patterns=True).fit(X[:32], y[:32])
    out['fit_coef'], out['fit_pattern'], out['fit_intercept'] = td.coef_, td.pattern_, td.intercept_
    out['fit_alpha'] = td.estimator.alpha_
    out['fit_pred'] = td.predict(X[32:])
    td2 = ref.estimators.TimeDelayed(-0.1, 0.06, 50, alphas=alphas, alpha_per_target=True, fit_intercept=False).fit(X[:32], y[:32])
    out['fit2_coef'], out['fit2_alpha'], out['fit2_pred'] = td2.coef_, td2.estimator.alpha_, td2.predict(X[32:])
    # hierarchical: identity groups (15 sets) and explicit groups (7 sets)
    h, hs = ref.model_selection.hierarchical_score(pipe(), X, y, cv=5)
    out['hier_score'], out['hier_mask'] = hs, torch.stack(list(h.mask_))
    out['hier_sets'] = np.array([list(s) + [-1] * (4 - len(s)) for s in h.set_])
    groups = [[1, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]
    h2, hs2 = ref.model_selection.hierarchical_score(pipe(), X, y, groups=groups, cv=5)
    out['hier2_score'], out['hier2_mask'], out['groups'] = hs2, torch.stack(list(h2.mask_)), np.array(groups)
    out['hier2_sets'] = np.array([list(s) + [-1] * (3 - len(s)) for s in h2.set_])
    torch.manual_seed(1)
    s, ss = ref.model_selection.shapley_score(pipe(), X, y, groups=groups, cv=5, n_permutations=3)
    out['shap_score'], out['shap_order'] = ss, s.order_
    save('trf_small.npz', **out)


# ---------------------------------------------------------------------------------------- decoder
def gen_decoder():
    g = torch.Generator().manual_seed(21)
    N, C, T, F = 80, 12, 6, 3
    X = torch.randn(N, C, T, generator=g)
    B = torch.randn(C, F, generator=g)
    y = torch.einsum('nct,cf->nft', X, B) * torch.hann_window(T + 2)[1:-1] + 2.0 * torch.randn(N, F, T, generator=g)
    alphas = torch.logspace(-5, 5, 20)
    out = dict(X=X, y=y, alphas=alphas)
    coefs, alps, icpts, pats, preds = [], [], [], [], []
    for t in range(T):
        Xt, yt = X[..., t], y[..., t]
        m = _RidgeCV_torch(alphas=alphas, fit_intercept=True, alpha_per_target=True).fit(Xt, yt)
        # the pattern formula of ridgedecoder.py:267-285 evaluated with reference torch ops
        Xc, yc = Xt - Xt.mean(0, keepdim=True), yt - yt.mean(0, keepdim=True)
        pat = torch.cov(Xc.T).mm(m.coef_.T).mm(torch.linalg.pinv(torch.cov(yc.T)))
        coefs.append(m.coef_); alps.append(m.alpha_); icpts.append(m.intercept_); pats.append(pat)
        preds.append(m.predict(Xt))
    out.update(coef=torch.stack(coefs), alpha=torch.stack(alps), intercept=torch.stack(icpts),
               pattern=torch.stack(pats), pred=torch.stack(preds, -1))
    # Sliding over the last dim with the (working) RidgeCV estimator, through the reference class
    sl = ref.estimators.Sliding(ref.estimators.RidgeCV(alphas=alphas, alpha_per_target=True), dims=-1).fit(X, y)
    out['sliding_coef'] = sl.collect('coef_')
    out['sliding_pred'] = sl.predict(X)
    # single target column
    m1 = _RidgeCV_torch(alphas=alphas, alpha_per_target=True).fit(X[..., 0], y[:, 0, 0])
    out['single_coef'], out['single_alpha'] = m1.coef_, m1.alpha_
    save('decoder.npz', **out)


def gen_b2b():
    """B2B composed from the reference's working parts exactly as b2b.py:267-306 spells it out (the class itself
    cannot be constructed at this commit: it forwards `normalise` to RidgeCV)."""
    from mvpy.preprocessing.scaler import _Scaler_torch
    g = torch.Generator().manual_seed(31)
    N, C, F = 160, 14, 4
    E = torch.randn(N, F, generator=g)
    E[:, 1] = 0.7 * E[:, 0] + 0.3 * E[:, 1]                        # correlated targets; target 3 has no effect on X
    Fm = torch.randn(F, C, generator=g); Fm[3] = 0.0
    X = E @ Fm + torch.randn(N, C, generator=g)
    alphas = torch.logspace(-3, 3, 13)
    out = dict(X=X, y=E, alphas=alphas)
    for tag, per_target, norm in (('a', False, True), ('b', True, False)):
        half = N // 2
        dec = _RidgeCV_torch(alphas=alphas, fit_intercept=True, alpha_per_target=per_target).fit(X[:half], E[:half])
        sc = _Scaler_torch(with_mean=norm, with_std=norm).fit(dec.predict(X[half:]))
        enc = _RidgeCV_torch(alphas=alphas, fit_intercept=True, alpha_per_target=per_target).fit(sc.transform(dec.predict(X[half:])), E[half:])
        Xc, yc = X[:half] - X[:half].mean(0, keepdim=True), E[:half] - E[:half].mean(0, keepdim=True)
        out[f'{tag}_pattern'] = torch.cov(Xc.T).mm(dec.coef_.T).mm(torch.linalg.pinv(torch.cov(yc.T)))
        out[f'{tag}_causal'] = enc.coef_.diagonal()
        out[f'{tag}_dec_alpha'], out[f'{tag}_enc_alpha'] = dec.alpha_, enc.alpha_
        out[f'{tag}_pred'] = enc.predict(sc.transform(dec.predict(X)))
    save('b2b.npz', **out)


def gen_clf():
    """Classification pieces.  LabelBinariser, math.accuracy, math.roc_auc and the Roc_auc metric run through the
    reference; RidgeClassifier (cannot be constructed at this commit: `normalise` keyword) is composed from the
    reference's binariser + _RidgeCV_torch exactly as ridgeclassifier.py:356-470 and classifier.py:731-870 do."""
    from mvpy.preprocessing.labelbinariser import _LabelBinariser_torch
    g = torch.Generator().manual_seed(51)
    N, C = 150, 10
    X = torch.randn(N, C, generator=g)
    W = torch.randn(C, 3, generator=g)
    y0 = torch.argmax(X @ W + 0.8 * torch.randn(N, 3, generator=g), dim=1).float()              # three classes 0, 1, 2
    y1 = torch.where(X[:, 0] + 0.7 * torch.randn(N, generator=g) > 0, 9.0, 5.0)                  # two classes 5, 9
    y = torch.stack([y0, y1], dim=1)
    alphas = torch.logspace(-2, 3, 6)
    out = dict(X=X, y=y, alphas=alphas)
    lb = _LabelBinariser_torch(neg_label=-1.0, pos_label=1.0)
    L = lb.fit_transform(y)
    out['L'], out['L01'] = L, ref.preprocessing.LabelBinariser().fit_transform(y)
    out['L_inv'] = lb.inverse_transform(L)
    m = _RidgeCV_torch(alphas=alphas, fit_intercept=True, alpha_per_target=False).fit(X, L)
    df = m.predict(X)
    onehot = torch.zeros_like(df)
    for s, e in ((0, 3), (3, 5)):
        onehot[torch.arange(N), s + torch.argmax(df[:, s:e], dim=1)] = 1.0
    pred = lb.inverse_transform(onehot)
    proba = 1 / (1 + torch.exp(-df))
    for s, e in ((0, 3), (3, 5)):
        proba[:, s:e] = proba[:, s:e] / proba[:, s:e].sum(-1, keepdim=True)
    Xc, Lc = X - X.mean(0, keepdim=True), L - L.mean(0, keepdim=True)
    out.update(ovr_alpha=m.alpha_, ovr_coef=m.coef_, ovr_intercept=m.intercept_, ovr_df=df, ovr_pred=pred, ovr_proba=proba,
               ovr_pattern=torch.cov(Xc.T).mm(m.coef_.T).mm(torch.linalg.pinv(torch.cov(Lc.T))))
    out['ovr_accuracy'] = ref.math.accuracy(y.T, pred.T)
    out['ovr_roc_auc'] = ref.metrics.roc_auc(y.T, df.T)
    # one-versus-one on both label features: pairwise fits on the two classes' samples, majority vote
    votes_all, ovo_coef = [], []
    cols = []
    for i, labels in enumerate((torch.tensor([0.0, 1.0, 2.0]), torch.tensor([5.0, 9.0]))):
        votes = []
        for j in range(len(labels)):
            for k in range(len(labels)):
                if j <= k:
                    continue
                sel = (y[:, i] == labels[j]) | (y[:, i] == labels[k])
                lb2 = _LabelBinariser_torch(neg_label=-1.0, pos_label=1.0)
                L2 = lb2.fit_transform(y[sel, i][:, None])
                m2 = _RidgeCV_torch(alphas=alphas, fit_intercept=True, alpha_per_target=False).fit(X[sel], L2)
                d2 = m2.predict(X)
                oh = torch.zeros_like(d2)
                oh[torch.arange(N), torch.argmax(d2, dim=1)] = 1.0
                votes.append(lb2.inverse_transform(oh).squeeze())
                ovo_coef.append(m2.coef_)
        votes = torch.stack(votes)
        counts = torch.stack([(votes == l).float().sum(0) for l in labels], dim=1)
        cols.append(labels[torch.argmax(counts, dim=1)])
    out['ovo_pred'], out['ovo_coef'] = torch.stack(cols).T, torch.stack(ovo_coef)
    # metric functions on their own: ties, multi-class one-versus-rest, a time axis, undefined cells
    s = torch.randn(4, 3, 60, generator=g)
    s[0, 0, :10] = s[0, 0, 10:20]                                                 # tied scores
    t2 = (torch.randn(4, 3, 60, generator=g) + 0.8 * s > 0).float() * 2 - 1
    t2[1, 1] = 1.0                                                                # one class only -> nan (kept by unique over the whole tensor)
    out.update(auc_s=s, auc_t2=t2, auc_bin=ref.math.roc_auc(t2, s))
    t3 = torch.randint(0, 3, (5, 80), generator=g).float()
    s3 = torch.randn(5, 3, 80, generator=g) + torch.nn.functional.one_hot(t3.long(), 3).float().swapaxes(-1, -2)
    out.update(auc_t3=t3, auc_s3=s3, auc_multi=ref.math.roc_auc(t3, s3))
    a, b = torch.randint(0, 3, (6, 7, 40), generator=g).float(), torch.randint(0, 3, (6, 7, 40), generator=g).float()
    out.update(acc_a=a, acc_b=b, acc=ref.math.accuracy(a, b))
    yt = torch.stack([y0, y1], dim=0)[:, None, :].expand(-1, 4, -1).contiguous()                 # (features, time, samples)
    dft = df.T[:, None, :] + 0.3 * torch.randn(5, 4, N, generator=g)
    out.update(auc_metric_y=yt, auc_metric_df=dft, auc_metric=ref.metrics.roc_auc(yt, dft))
    save('clf.npz', **out)


def rf_problem(seed, N, C, F, T, s_lo, s_hi, noise=1.0):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(N, C, T, generator=g, dtype=torch.float64)
    X = X + 0.6 * torch.roll(X, 1, -1) + 0.5                       # coloured, non-zero mean
    w = torch.randn(F, C, s_hi - s_lo, generator=g, dtype=torch.float64)
    y = torch.zeros(N, F, T, dtype=torch.float64)
    for k in range(s_hi - s_lo):
        lag = s_lo + k
        Z = torch.zeros_like(X)
        if lag >= 0:
            Z[..., lag:] = X[..., :T - lag]
        else:
            Z[..., :T + lag] = X[..., -lag:]
        y += torch.einsum('fc,nct->nft', w[:, :, k], Z)
    y = y + noise * y.std() * torch.randn(N, F, T, generator=g, dtype=torch.float64) + 1.5
    return X, y


RF_CASES = {
    # name: (problem args, (t_min, t_max, fs), constructor kwargs)
    'single': ((41, 10, 3, 2, 120, -4, 7), (-0.04, 0.06, 100), dict(alpha=1.0)),
    'kfold': ((41, 10, 3, 2, 120, -4, 7), (-0.04, 0.06, 100), dict(alpha=torch.logspace(-1, 5, 7), reg_cv=5)),
    'loo': ((41, 10, 3, 2, 120, -4, 7), (-0.04, 0.06, 100), dict(alpha=torch.logspace(-1, 5, 7), reg_cv='LOO')),
    'laplacian': ((42, 9, 3, 2, 100, -3, 5), (-0.03, 0.04, 100), dict(alpha=[0.1, 10.0, 1000.0], reg_type='laplacian', reg_cv=3, patterns=True)),
    'noedge': ((43, 6, 2, 3, 90, -5, 4), (-0.05, 0.03, 100), dict(alpha=2.0, edge_correction=False, fit_intercept=False, reg_type=['laplacian', 'ridge'])),
    'causal': ((44, 8, 1, 1, 150, 2, 9), (0.02, 0.08, 100), dict(alpha=0.5, patterns=True)),
    'anticausal': ((45, 8, 2, 2, 150, -8, -1), (-0.08, -0.02, 100), dict(alpha=3.0, reg_type=('ridge', 'laplacian'))),
}


RF_REG_CASES = [(3, 4, 'laplacian'), (3, 4, ('ridge', 'laplacian')), (3, 4, ('laplacian', 'ridge')), (1, 5, 'laplacian'),
                (4, 1, 'laplacian'), (2, 2, 'laplacian'), (3, 4, 'ridge'), (5, 3, ['laplacian', 'laplacian'])]


def gen_rf():
    """ReceptiveField through the reference's public class (torch.compile switched off by MVPY_COMPILE_TORCH=0: the
    decorated helper is plain Python then), fp64 inputs."""
    assert os.environ.get('MVPY_COMPILE_TORCH') == '0', 'run with MVPY_COMPILE_TORCH=0 (no C++ toolchain for inductor here)'
    out = {}
    for name, (prob, (t0, t1, fs), kw) in RF_CASES.items():
        X, y = rf_problem(*prob)
        if name == 'causal':
            X, y = X[:, 0], y[:, 0]                                # 2-D inputs: one feature, one channel
        m = ref.estimators.ReceptiveField(t0, t1, fs, **kw).fit(X, y)
        out[f'{name}_X'], out[f'{name}_y'] = X, y
        out[f'{name}_cov'], out[f'{name}_coef'], out[f'{name}_alpha'] = m.cov_, m.coef_, m.alpha_
        out[f'{name}_intercept'] = m.intercept_
        if m.pattern_ is not None:
            out[f'{name}_pattern'] = m.pattern_
        out[f'{name}_pred'] = m.predict(X)
        print(name, 's', m.s_min, m.s_max, 'alpha_', m.alpha_, 'cov asym', float((m.cov_ - m.cov_.transpose(1, 2)).abs().max()))
    from mvpy.estimators.receptivefield import _ReceptiveField_torch
    probe = _ReceptiveField_torch(-0.01, 0.01, 100)
    for i, (C, W, rt) in enumerate(RF_REG_CASES):
        out[f'reg{i}'] = probe._compute_reg_neighbours(C, W, rt)
    save('rf.npz', **out)


# ---------------------------------------------------------------------------------------- encoder
def gen_coupling():
    """One-versus-one probabilities.  The coupling routine itself runs through the reference (plain Python with
    MVPY_COMPILE_TORCH=0); the predict_proba flow around it (classifier.py:941-975) is composed from the reference's
    binariser + _RidgeCV_torch because RidgeClassifier cannot be constructed at this commit."""
    assert os.environ.get('MVPY_COMPILE_TORCH') == '0', 'run with MVPY_COMPILE_TORCH=0 (no C++ toolchain for inductor here)'
    from mvpy.estimators.classifier import _pairwise_coupling_torch
    from mvpy.preprocessing.labelbinariser import _LabelBinariser_torch
    g = torch.Generator().manual_seed(77)
    out = {}
    for tag, (N, K, dt) in {'a': (50, 3, torch.float64), 'b': (30, 5, torch.float64), 'c': (12, 2, torch.float64), 'd': (40, 4, torch.float32),
                            'e': (25, 7, torch.float32)}.items():
        j_i, k_i = torch.tril_indices(K, K, offset=-1)
        pind = (torch.rand(len(j_i), N, generator=g, dtype=torch.float64) * 0.96 + 0.02).to(dt)
        p_i = torch.zeros(N, K, K, dtype=dt)
        p_i[:, k_i, j_i] = pind.t()
        p_i[:, j_i, k_i] = (1 - pind).t()
        eps = 5e-3 / K
        out[f'{tag}_pair'] = p_i
        out[f'{tag}_p'] = _pairwise_coupling_torch(p_i, j_i, k_i, max_iter=max(100, K), eps=eps, tol=eps)
    c = np.load(os.path.join(OUT, 'clf.npz'))
    X, y, alphas = torch.from_numpy(c['X']), torch.from_numpy(c['y']), torch.from_numpy(c['alphas'])
    N = X.shape[0]
    probs = []
    for i, labels in enumerate((torch.tensor([0.0, 1.0, 2.0]), torch.tensor([5.0, 9.0]))):
        K = len(labels)
        dec = []
        for j in range(K):
            for k in range(K):
                if j <= k:
                    continue
                sel = (y[:, i] == labels[j]) | (y[:, i] == labels[k])
                L2 = _LabelBinariser_torch(neg_label=-1.0, pos_label=1.0).fit_transform(y[sel, i][:, None])
                dec.append(_RidgeCV_torch(alphas=alphas, fit_intercept=True, alpha_per_target=False).fit(X[sel], L2).predict(X))
        p_ind = 1 / (1 + torch.exp(-torch.stack(dec)))                       # (pairs, n, 2)
        j_i, k_i = torch.tril_indices(K, K, offset=-1)
        p_i = torch.zeros((N, K, K), dtype=X.dtype)
        p_i[:, k_i, j_i] = p_ind[:, :, 0].t()
        p_i[:, j_i, k_i] = p_ind[:, :, 1].t()
        eps = 5e-3 / K
        probs.append(_pairwise_coupling_torch(p_i, j_i, k_i, max_iter=max(100, K), eps=eps, tol=eps))
    out['ovo_proba'] = torch.cat(probs, dim=-1)
    save('coupling.npz', **out)


def gen_encoder():
    g = torch.Generator().manual_seed(31)
    N, F, T, D = 30, 4, 5, 3
    X = torch.randn(N, F, T, generator=g)
    y = torch.randn(N, D, T, generator=g) + torch.einsum('nft,fd->ndt', X, torch.randn(F, D, generator=g))
    alphas = torch.logspace(-3, 3, 9)
    enc = ref.estimators.RidgeEncoder(alphas=alphas).fit(X, y)
    out = dict(X=X, y=y, alphas=alphas, coef3=enc.coef_, intercept3=enc.intercept_, alpha3=enc.estimator.alpha_,
               pred3=enc.predict(X))
    enc2 = ref.estimators.RidgeEncoder(alphas=alphas).fit(X[..., 0], y[..., 0])
    out.update(coef2=enc2.coef_, intercept2=enc2.intercept_, alpha2=enc2.estimator.alpha_, pred2=enc2.predict(X[..., 0]))
    save('encoder.npz', **out)


# ------------------------------------------------------------------------------------------- cfg1
def gen_cfg1():
    torch.manual_seed(0)
    X, y = ref.dataset.make_meeg_continuous(fs=200)
    alphas = torch.logspace(-5, 5, 20)
    pipe = make_pipeline(ref.preprocessing.Scaler().to_torch(), ref.estimators.TimeDelayed(-1.0, 0.0, 200, alphas=alphas))
    val, sc = ref.crossvalidation.cross_val_score(pipe, X, y, cv=5)
    coef = torch.stack([m[-1].coef_ for m in val.model_])
    save('cfg1_summary.npz', alpha=torch.stack([m[-1].estimator.alpha_ for m in val.model_]),
         score_mean=sc.mean(), score_fold_mean=sc.mean((1, 2)), score_sub=sc[:, ::7, ::13],
         coef_sub=coef[:, ::7, :, ::5], intercept=torch.stack([m[-1].intercept_ for m in val.model_]),
         x_checksum=X.double().sum(), y_checksum=y.double().sum())


if __name__ == '__main__':
    which = sys.argv[1:] or ['ridgecv', 'delay', 'scaler', 'kfold', 'splitters', 'stratified', 'metrics', 'rank', 'trf_small', 'decoder', 'b2b', 'rf', 'clf', 'coupling', 'encoder', 'cfg1']
    for w in which:
        globals()['gen_' + w]()
