"""Host-side logic that needs no GPU: fold indices, lag windows, group enumeration, the metric protocol,
driver argument handling.  Index-type outputs are compared bit-exactly with the reference's goldens."""
import numpy as np
import pytest
import torch

import mvpy_b200 as mv
from mvpy_b200.estimators.timedelayed import lag_window
from mvpy_b200.model_selection.hierarchical import check_dims_and_groups_

T = torch.from_numpy


def test_kfold_bit_exact(golden):
    g = golden('kfold.npz')
    for n, k in [(10, 5), (11, 3), (120, 5), (7, 7)]:
        for i, (tr, te) in enumerate(mv.crossvalidation.KFold(n_splits=k).split(torch.zeros(n, 2))):
            assert tr.dtype == torch.long and np.array_equal(tr.numpy(), g[f'n{n}k{k}_tr{i}'])
            assert np.array_equal(te.numpy(), g[f'n{n}k{k}_te{i}'])
    for n, k, seed in [(20, 4, 0), (23, 5, 123)]:
        for i, (tr, te) in enumerate(mv.crossvalidation.KFold(n_splits=k, shuffle=True, random_state=seed).split(torch.zeros(n, 2))):
            assert np.array_equal(tr.numpy(), g[f'n{n}k{k}s{seed}_tr{i}']) and np.array_equal(te.numpy(), g[f'n{n}k{k}s{seed}_te{i}'])
    with pytest.raises(ValueError):
        list(mv.crossvalidation.KFold(n_splits=5).split(torch.zeros(3, 2)))


def test_lag_windows(golden):
    g = golden('delay.npz')
    assert np.array_equal(lag_window(-0.03, 0.02, 100), g['window'])
    assert np.array_equal(lag_window(0.01, 0.03, 100), g['window2'])
    assert np.array_equal(lag_window(-1.0, 0.0, 200), g['window_cfg1'])
    td = mv.estimators.TimeDelayed(-1.0, 0.0, 200, alphas=torch.logspace(-5, 5, 20))
    assert td.window.dtype == torch.int32 and td.window.shape == (201,) and int(td.window[0]) == -200


def test_group_enumeration(golden):
    g = golden('trf_small.npz')
    X, y = torch.zeros(5, 4, 10), torch.zeros(5, 2, 10)
    dim, ids, sets, masks, groups = check_dims_and_groups_(X, y)
    assert dim == -2 and ids == [0, 1, 2, 3]
    assert np.array_equal(masks.numpy(), g['hier_mask'])
    assert [list(s) + [-1] * (4 - len(s)) for s in sets] == g['hier_sets'].tolist()
    _, _, sets2, masks2, _ = check_dims_and_groups_(X, y, groups=g['groups'].tolist())
    assert np.array_equal(masks2.numpy(), g['hier2_mask'])
    assert [list(s) + [-1] * (3 - len(s)) for s in sets2] == g['hier2_sets'].tolist()
    # cfg3 groups: baseline block in every row -> exactly 8 unique masks
    gr = torch.zeros(4, 64, dtype=torch.long)
    gr[:, :16] = 1
    for r in range(1, 4):
        gr[r, 16 * r:16 * (r + 1)] = 1
    _, _, sets3, masks3, _ = check_dims_and_groups_(torch.zeros(2, 64, 3), torch.zeros(2, 1, 3), groups=gr)
    assert len(sets3) == 8 and sets3 == [(0,), (1,), (2,), (3,), (1, 2), (1, 3), (2, 3), (1, 2, 3)]
    assert masks3.sum(1).tolist() == [16, 32, 32, 32, 48, 48, 48, 64]
    with pytest.raises(ValueError):
        check_dims_and_groups_(X, y, groups=[[1, 0, 0, 0]])
    with pytest.raises(ValueError):
        check_dims_and_groups_(X, y, groups=[[1, 0, 0]])


def test_metric_protocol():
    m = mv.metrics.r2
    assert m.name == 'r2' and m.request == ('y', 'predict', 'y_b') and m.reduce == (0,)
    assert mv.metrics.pearsonr.request == ('y', 'predict')
    m2 = m.mutate(reduce=(2,))
    assert m2.reduce == (2,) and m.reduce == (0,) and m2.f is m.f
    m3 = mv.metrics.Metric()
    m3.grant({'a': 1})
    m3.grant({'b': 2})
    assert m3.metadata == {'a': 1, 'b': 2}
    with pytest.raises(ValueError):
        m3.grant([1])


def test_reduce_is_a_view_for_leading_dims():
    x = torch.arange(2 * 3 * 4.0).reshape(2, 3, 4)
    r = mv.metrics.reduce_(x, (0,))
    assert r.shape == (3, 4, 2) and r.data_ptr() == x.data_ptr()
    assert torch.equal(r[1, 2], x[:, 1, 2])
    assert mv.metrics.reduce_(x, (0, 2)).shape == (3, 8)
    assert torch.equal(mv.metrics.reduce_(x, -1), x)
    with pytest.raises(ValueError):
        mv.metrics.reduce_(x, (0, 0))
    with pytest.raises(ValueError):
        mv.metrics.reduce_(x, (5,))


class _MeanModel:
    """tiny user model (CPU torch) to exercise the drivers without a GPU"""
    metric_ = mv.metrics.Metric(name='mse', request=('y', 'predict'), reduce=(0,),
                                f=lambda y, p: ((y - p) ** 2).mean(-1))

    def fit(self, X, y):
        self.m_ = y.mean(0, keepdim=True) + 0 * X.sum()
        return self

    def predict(self, X):
        return self.m_.expand(X.shape[0], *self.m_.shape[1:])

    def score(self, X, y):
        return mv.metrics.score(self, self.metric_, X, y)

    def clone(self):
        return _MeanModel()


def test_validator_driver_on_host_model():
    X, y = torch.randn(20, 3), torch.randn(20, 2)
    val, sc = mv.crossvalidation.cross_val_score(_MeanModel(), X, y, cv=4)
    assert sc.shape == (4, 2) and len(val.model_) == 4 and len(val.test_) == 4
    for k, (tr, te) in enumerate(mv.crossvalidation.KFold(4).split(X)):
        assert torch.equal(val.test_[k], te)
        np.testing.assert_allclose(sc[k].numpy(), ((y[te] - y[tr].mean(0)) ** 2).mean(0).numpy(), rtol=1e-6)
    assert val.collect('m_').shape == (4, 1, 2)
    with pytest.raises(ValueError):
        mv.crossvalidation.Validator(object(), cv=3)
    with pytest.raises(ValueError):
        mv.crossvalidation.Validator(_MeanModel(), cv='nope')
    with pytest.raises(ValueError):
        mv.crossvalidation.Validator(_MeanModel()).collect('m_')
    # a splitter with n_repeats multiplies cv_n_ (validator.py:310-313)
    class Rep:
        n_splits, n_repeats = 3, 2
        def split(self, X, y=None):
            return iter(())
    assert mv.crossvalidation.Validator(_MeanModel(), cv=Rep()).cv_n_ == 6


def test_hierarchical_and_shapley_drivers_on_host_model():
    torch.manual_seed(3)
    X, y = torch.randn(12, 4, 5), torch.randn(12, 2, 5)
    h, hs = mv.model_selection.hierarchical_score(_MeanModel(), X, y, cv=3)
    assert hs.shape == (15, 3, 2, 5) and len(h.set_) == 15 and h.group_.shape == (4, 4)
    torch.manual_seed(5)
    s, ss = mv.model_selection.shapley_score(_MeanModel(), X, y, groups=[[1, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], cv=3,
                                             n_permutations=4)
    assert ss.shape == (4, 3, 3, 2, 5) and s.order_.shape == (4, 2)
    torch.manual_seed(5)                                     # RNG call order of the reference: perm then data, per permutation
    for i in range(4):
        perm = 1 + torch.randperm(2)
        torch.randperm(12)
        assert torch.equal(s.order_[i], perm)


def test_sliding_and_dispatch_arguments():
    est = mv.estimators.RidgeDecoder(alphas=torch.logspace(-1, 1, 3))
    assert mv.estimators.Sliding(est, dims=-1).dims == (-1,)
    assert mv.estimators.Sliding(est, dims=[1, 2]).dims == (1, 2)
    assert mv.estimators.Sliding(est, dims=np.array([2])).dims == (2,)
    with pytest.raises(ValueError):
        mv.estimators.Sliding(est, dims='x')
    with pytest.raises(ValueError):
        mv.estimators.Sliding(3)
    with pytest.raises(ValueError):
        mv.estimators.RidgeCV(alphas=np.array([1.0]))
    d = mv.estimators.RidgeDecoder(alphas=1.0, normalise=False)          # the keyword the reference trips over
    assert d.alpha_per_target is True and d.fit_intercept is True
    r = mv.estimators.RidgeCV(alphas=[0.1, 1.0])
    assert isinstance(r.alphas, torch.Tensor) and r.clone().alpha_per_target is False
    td = mv.estimators.TimeDelayed(-0.1, 0.0, 100, alphas=2.0, alpha_per_target=True)
    assert td.estimator.alpha_per_target is True and td.clone().kwargs == {'alpha_per_target': True}


def test_dataset_matches_reference_checksum(golden):
    g = golden('cfg1_summary.npz')
    torch.manual_seed(0)
    X, y = mv.dataset.make_meeg_continuous(fs=200)
    assert X.shape == (120, 3, 400) and y.shape == (120, 64, 400)
    assert abs(float(X.double().sum()) - float(g['x_checksum'])) < 1e-9
    assert abs(float(y.double().sum()) - float(g['y_checksum'])) < 1e-9


def test_repeated_kfold_and_leave_groups_out_bit_exact(golden):
    """Index listings of the reference's RepeatedKFold / LeaveGroupsOut (tests/golden/splitters.npz)."""
    g = golden('splitters.npz')
    rkf = mv.crossvalidation.RepeatedKFold(n_splits=3, n_repeats=2, random_state=7)
    folds = list(rkf.split(torch.zeros(11, 2)))
    assert len(folds) == 6
    for i, (tr, te) in enumerate(folds):
        assert tr.dtype == torch.long and np.array_equal(tr.numpy(), g[f'rkf_tr{i}']) and np.array_equal(te.numpy(), g[f'rkf_te{i}'])
    obj = mv.crossvalidation.RepeatedKFold(n_splits=2, n_repeats=2, random_state=1).to_torch()
    for rep in range(2):                       # a kept splitter continues its random stream across split() calls
        for i, (tr, te) in enumerate(obj.split(torch.zeros(6))):
            assert np.array_equal(tr.numpy(), g[f'rkfobj{rep}_tr{i}']) and np.array_equal(te.numpy(), g[f'rkfobj{rep}_te{i}'])
    lgo = mv.crossvalidation.LeaveGroupsOut(n_splits=2, n_repeats=2, random_state=3)
    for i, (tr, te) in enumerate(lgo.split(torch.zeros(12, 2), groups=T(g['g1']))):
        assert np.array_equal(tr.numpy(), g[f'lgo1_tr{i}']) and np.array_equal(te.numpy(), g[f'lgo1_te{i}'])
    lgo2 = mv.crossvalidation.LeaveGroupsOut(n_splits=2, n_repeats=1, random_state=5)
    seen = 0
    for i, (tr, te) in enumerate(lgo2.split(torch.zeros(12, 2), groups=T(g['g2']))):
        assert np.array_equal(tr.numpy(), g[f'lgo2_tr{i}']) and np.array_equal(te.numpy(), g[f'lgo2_te{i}'])
        assert len(tr) + len(te) < 12           # multi-label samples with labels on both sides are in neither
        seen += 1
    assert seen == 2
    with pytest.raises(ValueError):
        list(lgo.split(torch.zeros(12, 2)))
    # Validator accepts them: n_splits * n_repeats folds
    class M:
        metric_ = None
        def clone(self): return M()
        def fit(self, X, y): return self
        def score(self, X, y): return X.sum().reshape(1)
    v = mv.crossvalidation.Validator(M(), cv=mv.crossvalidation.RepeatedKFold(n_splits=3, n_repeats=2, random_state=0))
    assert v.cv_n_ == 6


def test_stratified_kfold_bit_exact(golden):
    """Index listings of the reference's StratifiedKFold / RepeatedStratifiedKFold (tests/golden/stratified.npz):
    single-feature, multi-feature and 3-D labels; shuffled folds draw from torch's global generator like the reference."""
    g = golden('stratified.npz')
    for name, k in (('s1', 3), ('s2', 4), ('s3', 4)):
        y = T(g['y' + name[1]])
        folds = list(mv.crossvalidation.StratifiedKFold(n_splits=k).split(torch.zeros(y.shape[0], 2), y))
        assert len(folds) == k
        for i, (tr, te) in enumerate(folds):
            assert tr.dtype == torch.long and np.array_equal(tr.numpy(), g[f'{name}_tr{i}']) and np.array_equal(te.numpy(), g[f'{name}_te{i}'])
    y2, y1 = T(g['y2']), T(g['y1'])
    torch.manual_seed(21)
    for i, (tr, te) in enumerate(mv.crossvalidation.StratifiedKFold(n_splits=4, shuffle=True, random_state=3).split(torch.zeros(40, 2), y2)):
        assert np.array_equal(tr.numpy(), g[f'sh_tr{i}']) and np.array_equal(te.numpy(), g[f'sh_te{i}'])
        # stratification: every joint label keeps its share in the held-out fold (within one sample)
        joint = y2[:, 0] * 100 + y2[:, 1]
        for v in joint.unique():
            assert abs(int((joint[te] == v).sum()) - int((joint == v).sum()) / 4) < 1.0
    torch.manual_seed(22)
    rsk = mv.crossvalidation.RepeatedStratifiedKFold(n_splits=3, n_repeats=2, random_state=5)
    folds = list(rsk.split(torch.zeros(17, 2), y1))
    assert len(folds) == 6 and rsk.n_splits * rsk.n_repeats == 6
    for i, (tr, te) in enumerate(folds):
        assert np.array_equal(tr.numpy(), g[f'rsk_tr{i}']) and np.array_equal(te.numpy(), g[f'rsk_te{i}'])
    with pytest.raises(ValueError):
        list(mv.crossvalidation.StratifiedKFold(n_splits=5).split(torch.zeros(3, 2), torch.zeros(3)))
    with pytest.raises(ValueError):          # every class smaller than n_splits
        list(mv.crossvalidation.StratifiedKFold(n_splits=4).split(torch.zeros(6, 2), torch.tensor([0, 0, 0, 1, 1, 1])))


def test_receptive_field_host_side(golden):
    """ReceptiveField pieces that need no device: lag range, argument checks, dispatch, penalty matrices
    (bit-equal to the reference's, tests/golden/rf.npz)."""
    from tests.rf_cases import RF_REG_CASES
    g = golden('rf.npz')
    rf = mv.estimators.ReceptiveField(-0.04, 0.06, 100, alpha=[1.0])
    assert (rf.s_min, rf.s_max) == (-4, 7) and rf.alpha == 1.0 and rf.window.tolist() == list(range(-4, 7))
    assert rf.metric_ is mv.metrics.r2 and rf.coef_ is None
    for i, (C, W, rt) in enumerate(RF_REG_CASES):
        assert np.array_equal(rf._compute_reg_neighbours(C, W, rt).numpy(), g[f'reg{i}']), (C, W, rt)
    for bad in (dict(t_min=0.1, t_max=0.0), dict(reg_type='laplacian', reg_cv='LOO'), dict(reg_cv=2.5), dict(alpha=None)):
        kw = dict(t_min=-0.1, t_max=0.1, fs=100); kw.update(bad)
        with pytest.raises(ValueError):
            mv.estimators.ReceptiveField(**kw)
    with pytest.raises(ValueError):
        rf._compute_reg_neighbours(2, 3, 'lasso')
    with pytest.raises(ValueError):
        rf._compute_reg_neighbours(2, 3, ('ridge',))
    with pytest.raises(ValueError):
        rf.predict(torch.zeros(2, 3, 10))
    c = mv.estimators.ReceptiveField(0.0, 0.1, 50, alpha=torch.logspace(-1, 1, 3), reg_cv=4, patterns=True).clone()
    assert c.reg_cv == 4 and c.patterns and torch.equal(c.alpha, torch.logspace(-1, 1, 3))
    x = torch.arange(5.0)[None]
    sh = type(rf)._shift
    assert sh(x, 2, 8).tolist() == [[0, 0, 0, 1, 2, 3, 4, 0]] and sh(x, -2, 4).tolist() == [[2, 3, 4, 0]] and sh(x, 9, 3).tolist() == [[0, 0, 0]]


def test_label_binariser_golden(golden):
    """LabelBinariser against the reference (tests/golden/clf.npz): two label features (3 and 2 classes), -1/+1 and 0/1
    codes, inverse transform, dtype following neg_label, error cases."""
    g = golden('clf.npz')
    y = T(g['y'])
    lb = mv.preprocessing.LabelBinariser(neg_label=-1.0, pos_label=1.0).to_torch()
    L = lb.fit_transform(y)
    assert L.dtype == torch.float32 and np.array_equal(L.numpy(), g['L'])
    assert np.array_equal(lb.inverse_transform(L).numpy(), g['L_inv']) and np.array_equal(g['L_inv'], g['y'])
    L01 = mv.preprocessing.LabelBinariser().fit_transform(y)
    assert L01.dtype == torch.int64 and np.array_equal(L01.numpy(), g['L01'])
    assert lb.n_features_ == 2 and lb.n_classes_ == [3, 2] and lb.C_.tolist() == [0.0, 3.0] and float(lb.N_) == 5.0
    assert [l.tolist() for l in lb.labels_] == [[0.0, 1.0, 2.0], [5.0, 9.0]]
    with pytest.raises(ValueError):
        lb.transform(torch.tensor([[0.0, 7.0]]))                  # unknown label
    with pytest.raises(ValueError):
        lb.transform(torch.zeros(3, 3))                           # wrong number of label features
    with pytest.raises(ValueError):
        mv.preprocessing.LabelBinariser(1, 0).to_torch()
    with pytest.raises(ValueError):
        lb.clone().transform(y)                                   # not fitted
    c = mv.estimators.RidgeClassifier(alphas=[1.0, 10.0], method='OvO', alpha_per_target=True)
    assert c.method == 'OvO' and c.kwarguments['alpha_per_target'] and c.metric_ is mv.metrics.accuracy
    with pytest.raises(ValueError):
        mv.estimators.Classifier(mv.estimators.RidgeClassifier, method='OvA')
    with pytest.raises(ValueError):
        c.predict(torch.zeros(2, 3))


def test_fold_lanes_host_logic():
    """mvpy_b200/_streams.py without a GPU: CPU tensors run the tasks in order on the caller's thread; lane count limits."""
    import torch
    from mvpy_b200 import _streams
    order = []
    out = _streams.run_on_lanes([(lambda i=i: (order.append(i), i * i)[1]) for i in range(5)], torch.device('cpu'), lane_keys=[0, 1, 0, 1, 2])
    assert out == [0, 1, 4, 9, 16] and order == [0, 1, 2, 3, 4]
    try:
        _streams.set_fold_streams(100)
        assert _streams.n_fold_streams() == _streams.MAX_LANES
        _streams.set_fold_streams(0)
        assert _streams.n_fold_streams() == 1
    finally:
        _streams.set_fold_streams(None)
    assert _streams.n_fold_streams() >= 1


def test_batched_executor_declines_without_a_gpu_or_for_other_estimators():
    import torch
    import mvpy_b200 as mv
    from mvpy_b200 import _batched
    X, y = torch.randn(50, 4, 6), torch.randn(50, 2, 6)
    pairs = [(i, i) for i in range(6)]
    if not torch.cuda.is_available():
        assert _batched.fit_slices(mv.estimators.RidgeDecoder(alphas=torch.tensor([1.0])), X, y, pairs) is None
    assert _batched.fit_slices(object(), X, y, pairs) is None
    assert _batched.fit_slices(mv.estimators.RidgeDecoder(alphas=torch.tensor([1.0])), X.double(), y.double(), pairs) is None      # float64: general route
    assert _batched.predict_slices([object()] * 6, X) is None
