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
    """Stands in for a fitted Validator: score_ depends on the predictor mask and on the (possibly reshuffled) data."""
    metric = None

    def __init__(self, score):
        self.score_ = score


def _stub_fit_validator(validator, X, y, mask, dim, shared=None):
    m = mask.to(torch.float32)
    w = torch.arange(1, m.numel() + 1, dtype=torch.float32)
    trial_sig = (X[:, 0, 0] * torch.arange(1, X.shape[0] + 1, dtype=torch.float32)).sum()     # order of the trials matters
    base = (m * w).sum() + 1e-3 * trial_sig
    return _StubValidator(base + torch.arange(5 * 2 * 3, dtype=torch.float32).reshape(5, 2, 3) * 0.01)


def _patch():
    import mvpy_b200  # noqa: F401
    from mvpy_b200 import _lib
    from mvpy_b200.model_selection import hierarchical, shapley
    hierarchical.fit_validator_ = _stub_fit_validator
    shapley.fit_validator_ = _stub_fit_validator
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
    proto.clone = lambda: proto
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
