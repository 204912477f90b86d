"""2-GPU check of the sharded drivers over NCCL (run under torchrun on a multi-GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py
Every rank runs hierarchical_score and shapley_score sharded over the ranks, then the same calls locally
(sharding disabled), and checks that the gathered scores / permutation orders are identical."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from sklearn.pipeline import make_pipeline
import mvpy_b200 as mv
from mvpy_b200 import _dist

rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
torch.manual_seed(0)
X, y = mv.dataset.make_meeg_continuous(fs=50, n_trials=40, n_channels=8, n_features=4)
X, y = X.to(dev), y.to(dev)
alphas = torch.logspace(-5, 5, 20)
pipe = make_pipeline(mv.preprocessing.Scaler().to_torch(), mv.estimators.TimeDelayed(-0.4, 0.0, 50, alphas=alphas))
groups = torch.tensor([[1, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])

def run():
    _, hs = mv.model_selection.hierarchical_score(pipe, X, y, groups=groups, cv=5)
    torch.manual_seed(1 + 100 * (rank if _dist.sharding_active() else 0))      # only rank 0's draws may matter when sharded
    sh, ss = mv.model_selection.shapley_score(pipe, X, y, groups=groups, cv=5, n_permutations=5)
    return hs, ss, sh.order_

hs_d, ss_d, ord_d = run()                       # sharded: units round-robin over the ranks, one all-gather
with _dist.sharded_region():                    # local: every rank computes everything itself
    hs_l, ss_l, ord_l = run()
ok = torch.equal(hs_d, hs_l) and torch.equal(ss_d, ss_l) and torch.equal(ord_d, ord_l)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f'dist_check world={dist.get_world_size()} hierarchical {tuple(hs_d.shape)} shapley {tuple(ss_d.shape)} order {ord_d.tolist()} '
          f'sharded == local on all ranks: {bool(flag.item())}', flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
