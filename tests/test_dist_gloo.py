"""N > 1 path on CPU: world_size-2 ``gloo`` runs of the unit sharding (``mvpy_b200/_dist.py``) and of the Hierarchical /
Shapley drivers on top of it.  The CUDA fits are replaced by a deterministic host stub (the product has no CPU path),
so what is checked is exactly the multi-process logic: unit assignment, rank-0-authoritative RNG draws, the single
all-gather of score shards, and equality with the single-process result."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port() -> int:
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _StubValidator:
    """Stands in for a (fitted) Validator: score_ depends on the predictor mask and on the (possibly reshuffled) data."""
    metric = None
    model = None

    def __init__(self, score):
        from mvpy_b200.crossvalidation import KFold
        self.score_ = score
        self.cv_ = KFold(n_splits=5)

    def clone(self):
        return _StubValidator(None)


_FOLD_TABLE = torch.arange(5 * 2 * 3, dtype=torch.float32).reshape(5, 2, 3) * 0.01


def _base_score(X, feat_weight_sum):
    trial_sig = (X[:, 0, 0] * torch.arange(1, X.shape[0] + 1, dtype=torch.float32)).sum()     # order of the trials matters
    return feat_weight_sum + 1e-3 * trial_sig


def _stub_fit_validator(validator, X, y, mask, dim, shared=None, split_like=None):
    m = mask.to(torch.float32)
    w = torch.arange(1, m.numel() + 1, dtype=torch.float32)
    return _StubValidator(_base_score(X, (m * w).sum()) + _FOLD_TABLE)


def _stub_fit_model(model, train, test, X, y=None, metric=None, train_key=None):
    """One (set, fold) unit of the sharded Hierarchical: the same numbers as fold k of _stub_fit_validator, rebuilt from the
    context the driver hands to the estimators (parent array + selected feature indices)."""
    from mvpy_b200 import _lib
    par = _lib.fold_ctx.parent
    k = [0, 3, 6, 8, 10].index(int(test[0]))                      # KFold(5) on 12 trials: test blocks start here
    return dict(model=None, score_=_base_score(par['X'], (par['idx'].float() + 1).sum()) + _FOLD_TABLE[k], test=test, metric=None)


def _patch():
    import mvpy_b200  # noqa: F401
    from mvpy_b200 import _lib
    from mvpy_b200.model_selection import hierarchical, shapley
    from mvpy_b200.crossvalidation import validator
    hierarchical.fit_validator_ = _stub_fit_validator
    shapley.fit_validator_ = _stub_fit_validator
    validator.fit_model_ = _stub_fit_model
    _lib.stage = lambda t: t
    hierarchical._move_tensors = lambda v, dev: v
    shapley._move_tensors = lambda v, dev: v
    return hierarchical, shapley


def _data():
    g = torch.Generator().manual_seed(7)
    X = torch.randn(12, 6, 10, generator=g)
    y = torch.randn(12, 2, 10, generator=g)
    groups = torch.tensor([[1, 1, 0, 0, 0, 0], [0, 0, 1, 1, 0, 0], [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 1]])
    return X, y, groups


def _run_drivers(seed):
    hierarchical, shapley = _patch()
    X, y, groups = _data()
    proto = _StubValidator(None)
    h = hierarchical.Hierarchical(proto).fit(X, y, groups=groups)
    torch.manual_seed(seed)
    s = shapley.Shapley(proto, n_permutations=7).fit(X, y, groups=groups)
    return h.score_.clone(), [tuple(t) for t in h.set_], s.score_.clone(), s.order_.clone(), [v is not None for v in s.validator_]


def _worker(rank, world, port, out_path):
    if world == 1:                       # single-process reference, in a child so the stubs never leak into pytest
        torch.save(_run_drivers(seed=1), f'{out_path}.single')
        return
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from mvpy_b200 import _dist
        # --- unit sharding primitives ---
        n_units = 5
        mine = _dist.my_units(n_units)
        assert mine == list(range(rank, n_units, world))
        local = {u: torch.full((2, 3), float(u)) for u in mine}
        full = _dist.gather_units(local, n_units)
        assert len(full) == n_units and all(torch.equal(full[u], torch.full((2, 3), float(u))) for u in range(n_units))
        assert _dist.broadcast_object({'rank': rank}, src=0) == {'rank': 0}
        # arbitrary (cost-balanced) assignment: longest first, deterministic, and the zero-fill + sum gather is exact
        owner = _dist.lpt_assign([5.0, 1.0, 1.0, 4.0, 1.0, 1.0], world)
        assert owner == [0, 1, 0, 1, 1, 0]
        local = {u: torch.full((2, 3), 0.1 * u - 0.3) for u in range(6) if owner[u] == rank}
        full = _dist.gather_assigned(local, 6)
        assert all(torch.equal(full[u], torch.full((2, 3), 0.1 * u - 0.3)) for u in range(6))
        with _dist.sharded_region():
            assert not _dist.sharding_active()          # nested drivers run locally
        assert _dist.sharding_active()
        # --- drivers: rank 1 seeds its RNG differently on purpose; rank 0's draws are authoritative ---
        res = _run_drivers(seed=1 if rank == 0 else 999)
        torch.save(res, f'{out_path}.{rank}')
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_gloo_matches_single_process(tmp_path):
    port = _free_port()
    out = str(tmp_path / 'res')
    mp.spawn(_worker, args=(1, port, out), nprocs=1, join=True)
    single = torch.load(f'{out}.single')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    for rank in range(2):
        h_score, h_sets, s_score, s_order, owned = torch.load(f'{out}.{rank}')
        assert h_sets == single[1]
        assert torch.equal(h_score, single[0])            # every rank holds every set's scores after the gather
        assert torch.equal(s_order, single[3])            # permutation orders bit-exact with the single-process run
        assert torch.equal(s_score, single[2])
        assert owned == [(i % 2) == rank for i in range(7)]   # round-robin ownership of the permutations
