"""Flat spherical k-means with rows sharded over ranks and an NCCL all-reduce of the centroid sums per
iteration (BASELINE configs[2]/[4] mode), against the single-GPU run of the same problem: the exact int64
fixed-point sums make the labels bit-identical at every GPU count.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dist_kmeans.py [N] [D] [K] [T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from hsg_b200.models import utils as MU
from hsg_b200.utils.segsort import common as S

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 256
K = int(sys.argv[3]) if len(sys.argv) > 3 else 256
T = int(sys.argv[4]) if len(sys.argv) > 4 else 20
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
dist.init_process_group('nccl', device_id=dev)
g = torch.Generator(device=dev); g.manual_seed(235)
x = torch.randn(N, D, device=dev, generator=g)
x = x / x.norm(dim=1, keepdim=True)
init = torch.randint(0, K, (N,), device=dev, generator=g)
lo, hi = N * rank // world, N * (rank + 1) // world
MU.dist_kmeans_with_initial_labels(x[lo:hi], init[lo:hi], K, 2)
dist.barrier(); torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
mine = MU.dist_kmeans_with_initial_labels(x[lo:hi], init[lo:hi], K, T)
t1.record(); torch.cuda.synchronize()
ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
full = MU.dist_kmeans_with_initial_labels(x, init, K, T, collective=False)   # every rank alone: the whole problem on one GPU
agree = (mine == full[lo:hi]).float().mean()
dist.all_reduce(agree, op=dist.ReduceOp.SUM)
if rank == 0:
  print('flat k-means N=%d D=%d K=%d T=%d on %d GPUs (rows sharded, all-reduce of [K,D] sums per iteration): '
        '%.3f ms/iteration (max over ranks), labels equal to the single-GPU run on %.5f of the rows'
        % (N, D, K, T, world, float(ms) / T, float(agree) / world))
dist.destroy_process_group()
