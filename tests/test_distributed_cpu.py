"""N>1 host logic on CPU (gloo, world_size 2): stack sharding and the one-all-reduce gradient bucket with the reference's
global-batch loss normalisation."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dffinthewild_b200 import distributed as D
    from dffinthewild_b200.Depth_Estimation_Network import Network
    torch.manual_seed(0)
    net = Network()
    skip = D.unused_parameter_names(net)
    bucket = D.GradBucket(net, skip=skip)
    # synthetic per-rank gradients and per-rank valid-pixel counts
    g = torch.Generator().manual_seed(100 + rank)
    local = torch.randn(bucket.numel, generator=g)
    bucket.zero()
    bucket.flat[:bucket.numel].copy_(local)
    n_r = float(1000 + 500 * rank)
    bucket.allreduce_gradients(weight=n_r)
    # expected: sum_r n_r g_r / sum_r n_r
    num, den = torch.zeros(bucket.numel), 0.0
    for r in range(world):
        gr = torch.randn(bucket.numel, generator=torch.Generator().manual_seed(100 + r))
        num += (1000 + 500 * r) * gr
        den += 1000 + 500 * r
    err = (bucket.flat[:bucket.numel] - num / den).abs().max().item()
    first = next(p for n, p in net.named_parameters() if n not in skip)
    views_ok = first.grad.data_ptr() == bucket.flat.data_ptr()
    q.put((rank, err, bucket.numel, len(skip), views_ok, D.shard_batch(7, rank, world)))
    dist.destroy_process_group()


def test_gradient_bucket_allreduce_world2():
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, numel, nskip, views_ok, shard in res:
        assert err < 1e-5
        assert nskip == 12 and numel == 4038832 - 22240   # every parameter that receives a gradient
        assert views_ok
    assert res[0][5] == [0, 2, 4, 6] and res[1][5] == [1, 3, 5]
