"""Multi-GPU training step of the DefocusNet-shaped config (BASELINE.json configs[2]: 5x3x256x256, 4 stacks per GPU):
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/ddp_train.py [--steps K] [--check]
One process per GPU; per-rank BatchNorm statistics (as nn.DataParallel); ONE NCCL all-reduce of the flat gradient bucket per
step, weighted by the per-rank valid-pixel count (global-batch masked-MSE normalisation of train_code_Defocus.py:160-165).
--check: rank 0 recomputes every rank's gradient on its own GPU and verifies the all-reduced bucket equals the weighted mean."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from dffinthewild_b200 import distributed as D
from dffinthewild_b200.Depth_Estimation_Network import Network
from dffinthewild_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--check", action="store_true")
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, S, H, W = 4, 5, 256, 256
crit = torch.nn.MSELoss()


def make(seed_rank):
    torch.manual_seed(0)
    net = Network()
    net.load_state_dict(synth.synthetic_state(net.state_dict(), seed=1))
    net = net.to(dev).train()
    FS, fd = synth.focal_stack(B, S, H, W, seed=200 + seed_rank).to(dev), synth.focus_dists(B, S, H, W, "defocus").to(dev)
    gt, mask = synth.gt_and_mask(B, H, W, seed=200 + seed_rank)
    return net, FS, fd, gt.to(dev), mask.to(dev)


def loss_of(net, FS, fd, gt, mask):
    o = net(FS, fd)
    return 0.5 * crit(o[1][mask], gt[mask]) + 0.7 * crit(o[2][mask], gt[mask]) + crit(o[3][mask], gt[mask]) + 0.3 * crit(o[0][mask], gt[mask])


net, FS, fd, gt, mask = make(rank)
bucket = D.GradBucket(net, skip=D.unused_parameter_names(net))
opt = torch.optim.Adam(bucket.params, lr=1e-4, betas=(0.9, 0.99))
n_r = float(mask.sum().item())

if args.check:
    bucket.zero()
    loss_of(net, FS, fd, gt, mask).backward()
    bucket.allreduce_gradients(weight=n_r)
    got = bucket.flat[:bucket.numel].clone()
    if rank == 0:
        num, den = torch.zeros_like(got), 0.0
        for r in range(world):
            n2, FS2, fd2, gt2, mask2 = make(r)
            b2 = D.GradBucket(n2, skip=D.unused_parameter_names(n2))
            loss_of(n2, FS2, fd2, gt2, mask2).backward()
            w = float(mask2.sum().item())
            num += w * b2.flat[:b2.numel]
            den += w
        ref = num / den
        cos = float(torch.dot(ref.double(), got.double()) / (ref.double().norm() * got.double().norm()))
        print(json.dumps({"check": "allreduced gradient bucket vs weighted mean of per-rank gradients", "world": world,
                          "cosine": cos, "max_abs_diff": float((ref - got).abs().max()), "max_abs": float(ref.abs().max())}))
        assert cos > 0.99999

for i in range(args.warmup + args.steps):
    if i == args.warmup:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    bucket.zero()
    loss = loss_of(net, FS, fd, gt, mask)
    loss.backward()
    bucket.allreduce_gradients(weight=n_r)
    opt.step()
e1.record()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"metric": "DefocusNet-shape training focal stacks/sec (fwd + loss + bwd + all-reduce + Adam)", "value": B * world / (ms.item() / 1e3),
                      "unit": "stacks/s", "n_gpus": world, "ms_per_step": ms.item(), "scaling": "weak", "dtype": "f32",
                      "config": {"workload": "5x3x256x256, 4 stacks per GPU, fp32 parity path", "allreduce_bytes": 4 * (bucket.numel + 1)}}))
if world > 1:
    dist.destroy_process_group()
