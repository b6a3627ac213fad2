#!/usr/bin/env python
"""bench.py — focal stacks/sec of the depth-from-focus forward on DDFF-12-shaped stacks (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16]

One step = one pass of the hot path (DFF_net forward, reference train_codes/Depth_Estimation_Network.py:77-137)
over 64 synthetic DDFF full-resolution stacks PER GPU (10 x 3 x 383 x 552, padded to 384 x 576 with -1 like
Depth_Estimation_Test/test_Dataloader.py:128-140), one dff_forward call per step (--micro-batch 64; 49 GB of workspace).  Focal stacks are independent, so ranks share nothing: weak scaling,
no collective on the data path (N=1 is exactly BASELINE.json configs[1]: batch 64).

`value`   stacks/s, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`     the same metric through the C-ABI call that takes HOST buffers (`dff_forward_host`): pinned-host -> device
          copies of FS / focus_dists and device -> host reads of the four depth maps are inside the timed region.
`roofline` the dominant kernel (largest share of the step): algorithmic FLOPs / its CUDA-event time vs the measured
          bf16 tensor peak of MEASURED_PEAKS.json.
`cpu_baseline` the oracle port (the reference's torch CPU ops) on this box's host cores, one stack of the workload.
`--impl reference` times that CPU path alone (rank 0 only).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PER_GPU_BATCH, S, H, W, VALID_HW = 64, 10, 384, 576, (383, 552)
FLOP_PER_VOXEL = 93563.0   # 2*MACs over the 70 executed conv layers
METRIC = "DDFF-shape focal stacks/sec"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons, sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_net(precision):
    import torch
    from dffinthewild_b200.Depth_Estimation_Network import Network
    from dffinthewild_b200 import synth
    torch.manual_seed(0)
    net = Network()
    sd = synth.synthetic_state(net.state_dict(), seed=1)
    net.load_state_dict(sd)
    net.DFF_net.precision = precision
    return net, sd


def cpu_reference_time(sd, n_runs, warmup):
    """The reference's CPU implementation of the path (oracle port: same torch CPU ops in the same order) on one
    stack of the workload, all host threads."""
    import torch
    from oracle import dff_oracle
    from dffinthewild_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    FS = synth.focal_stack(1, S, H, W, seed=0, valid_hw=VALID_HW)
    fd = synth.focus_dists(1, S, H, W, "ddff")
    times = []
    with torch.no_grad():
        for i in range(warmup + n_runs):
            t0 = time.perf_counter()
            dff_oracle.dff_forward(sd, FS, fd)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return times, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _, sd = make_net("fp32")
    steps = max(1, args.steps)
    times, cores = cpu_reference_time(sd, steps, min(args.warmup, 1))
    total = sum(times)
    val = len(times) / total
    sample = "1 stack (10x3x384x576) per step, oracle port of the reference's torch CPU path, %d torch threads" % cores
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "stacks/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1000 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, 1, "cpu"),
        "cpu_baseline": {"value": val, "unit": "stacks/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "stacks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(n_gpus, micro_batch, precision):
    return {"workload": "DDFF-12 full-res inference (BASELINE.json configs[1]): 64 stacks of 10x3x383x552 (padded to 384x576) "
                        "per GPU, independent stacks partitioned across GPUs",
            "stacks_per_gpu": PER_GPU_BATCH, "global_batch": PER_GPU_BATCH * n_gpus, "slices": S, "padded_hw": [H, W],
            "micro_batch": micro_batch, "precision": precision,
            "parallelism": "stack-partitioned x%d, no collective" % n_gpus,
            "l2": "inputs (%.1f GB per rank) exceed the 126 MB L2; no explicit flush" % (PER_GPU_BATCH * 4 * 4 * S * H * W / 1e9)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from dffinthewild_b200 import runtime as rt
    from dffinthewild_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n_local = PER_GPU_BATCH
    mb = min(args.micro_batch, n_local)
    assert n_local % mb == 0
    net, sd = make_net(args.precision)
    net = net.to(dev).eval()
    dff = net.DFF_net
    lib = rt.lib()
    mode = rt.BF16 if args.precision == "bf16" else rt.FP32

    # ---- synthetic inputs, resident in HBM -----------------------------------------------------------------
    one = synth.focal_stack(mb, S, H, W, seed=100 + rank, valid_hw=VALID_HW)
    FS = torch.empty((n_local, 3, S, H, W), dtype=torch.float32, device=dev)
    for i in range(0, n_local, mb):
        FS[i:i + mb] = one.to(dev).roll(i, dims=-1)   # distinct content per chunk
    fd = synth.focus_dists(n_local, S, H, W, "ddff").to(dev)
    outs = [torch.empty((n_local, H, W), dtype=torch.float32, device=dev) for _ in range(4)]
    packed = rt.packed_weights(dff, dev)
    ws = torch.empty(lib.dff_workspace_bytes(mb, S, H, W, mode), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)
    strides = (ctypes.c_int64 * 4)(*fd[:mb].stride())

    def chunk_ptrs(i):
        return (ctypes.c_void_p * 4)(*[o[i:i + mb].data_ptr() for o in outs])

    def step():
        for i in range(0, n_local, mb):
            rt.check(lib.dff_forward(packed.data_ptr(), FS[i:i + mb].data_ptr(), fd[i:i + mb].data_ptr(), strides, mb, S, H, W,
                                     chunk_ptrs(i), None, ws.data_ptr(), ws.numel(), mode, local, sp))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = PER_GPU_BATCH * world / (ms_per_step / 1000.0)

    # ---- per-operator profile (CUDA events on the launching stream) -> dominant kernel + roofline ---------------
    NOPS = 256
    op_ms, op_fl, op_by = (ctypes.c_float * NOPS)(), (ctypes.c_double * NOPS)(), (ctypes.c_double * NOPS)()
    op_la, op_nm, n_ops = (ctypes.c_int * NOPS)(), ctypes.create_string_buffer(NOPS * 64), ctypes.c_int(0)
    agg = {}
    prof_chunks = min(3, n_local // mb)
    for j in range(prof_chunks):
        i = j * mb
        rt.check(lib.dff_forward_profiled(packed.data_ptr(), FS[i:i + mb].data_ptr(), fd[i:i + mb].data_ptr(), strides, mb, S, H,
                                          W, chunk_ptrs(i), ws.data_ptr(), ws.numel(), mode, local, sp, NOPS, op_ms, op_fl, op_by,
                                          op_la, op_nm, ctypes.byref(n_ops)))
        for k in range(n_ops.value):
            name = op_nm.raw[k * 64:(k + 1) * 64].split(b"\0")[0].decode()
            a = agg.setdefault(name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0, "calls": 0})
            a["ms"] += op_ms[k]; a["flops"] += op_fl[k]; a["bytes"] += op_by[k]; a["launches"] += op_la[k]; a["calls"] += 1
    launches_per_chunk = sum(op_la[k] for k in range(n_ops.value))
    total_prof_ms = sum(a["ms"] for a in agg.values())
    top_name, top = max(agg.items(), key=lambda kv: kv[1]["ms"])
    pk = peaks()
    ach_tf = top["flops"] / (top["ms"] / 1000.0) / 1e12
    ach_gbs = top["bytes"] / (top["ms"] / 1000.0) / 1e9
    roof = {"kernel": top_name, "bound": "tensor", "achieved": ach_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
            "frac": ach_tf / pk["tf_sustained"], "traffic": None, "peak_source": pk["src"] + " (sustained bf16 GEMM)",
            "share_of_step": top["ms"] / total_prof_ms, "avg_launch_ms": top["ms"] / max(top["launches"], 1),
            "algorithmic_flops_per_launch": top["flops"] / max(top["launches"], 1),
            "hbm_view": {"achieved_gbs": ach_gbs, "peak_gbs": pk["hbm"], "frac": ach_gbs / pk["hbm"]},
            "whole_step": {"tflops_per_gpu": FLOP_PER_VOXEL * S * H * W * PER_GPU_BATCH / (ms_per_step / 1000.0) / 1e12,
                           "frac_of_tensor_peak": FLOP_PER_VOXEL * S * H * W * PER_GPU_BATCH / (ms_per_step / 1000.0) / 1e12 / pk["tf_sustained"]}}
    try:   # BASELINE.json's second figure: tensor-pipe utilisation of the 3-D aggregation convs (hourglasses, pyramid, dres0, deconvs)
        sel = [a for n, a in agg.items() if n.startswith(("dres", "SPP_module", "deconv_", "confidence"))]
        t_s, f_s = sum(a["ms"] for a in sel) / 1000.0, sum(a["flops"] for a in sel)
        roof["aggregation_convs"] = {"tflops": f_s / t_s / 1e12, "frac_of_tensor_peak": f_s / t_s / 1e12 / pk["tf_sustained"],
                                     "share_of_step": 1000.0 * t_s / total_prof_ms, "operators": len(sel)}
    except Exception as ex:   # (a reporting extra must never take the bench line down)
        roof["aggregation_convs"] = {"error": str(ex)}
    tr = os.path.join(ROOT, "profiles", "top_kernel_traffic.json")   # dram bytes per launch of the dominant kernel, from the committed ncu --set full capture
    if os.path.exists(tr):
        t = json.load(open(tr))
        if t.get("kernel") == top_name and t.get("micro_batch") == mb and t.get("precision") == args.precision:
            roof["traffic"] = t["dram_bytes_per_launch"]
            roof["traffic_source"] = t.get("source")

    # ---- end to end through the host-buffer C-ABI call ------------------------------------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    hFS = torch.empty((n_local, 3, S, H, W), dtype=torch.float32).pin_memory()
    hFS.copy_(FS)
    hfd = torch.empty((n_local, S, H, W), dtype=torch.float32).pin_memory()
    hfd.copy_(fd)
    houts = [torch.empty((n_local, H, W), dtype=torch.float32).pin_memory() for _ in range(4)]
    emb = min(args.e2e_micro_batch, n_local)
    dev_io = torch.empty(lib.dff_host_io_bytes(emb, S, H, W), dtype=torch.uint8, device=dev)
    if emb > mb:
        ws = torch.empty(lib.dff_workspace_bytes(emb, S, H, W, mode), dtype=torch.uint8, device=dev)
    hstrides = (ctypes.c_int64 * 4)(*hfd.stride())
    hp = (ctypes.c_void_p * 4)(*[o.data_ptr() for o in houts])

    def e2e_step():
        # ONE C-ABI call per step: the library pipelines H2D copies / kernels / D2H reads over micro-batches internally
        rt.check(lib.dff_forward_host(packed.data_ptr(), hFS.data_ptr(), hfd.data_ptr(), hstrides, n_local, emb, S, H, W, hp,
                                      dev_io.data_ptr(), ws.data_ptr(), ws.numel(), mode, local, sp))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = PER_GPU_BATCH * world * e2e_steps / e2e_s
    h2d = n_local * (3 * S * H * W + S * H * W) * 4
    d2h = n_local * 4 * H * W * 4
    same = all(torch.equal(h.to(dev), o) for h, o in zip(houts, outs))

    line = {
        "metric": METRIC, "value": value, "unit": "stacks/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(world, mb, args.precision),
        "clocks": clk.summary(),
        "e2e": {"value": e2e_val, "unit": "stacks/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "dff_forward_host (C-ABI, pinned host buffers; copies pipelined with kernels over micro-batches of <= %d)" % emb,
                "matches_device_run": bool(same)},
        "gpu_launches": launches_per_chunk * (n_local // mb) * args.steps,
        "roofline": roof,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        times, cores = cpu_reference_time(sd, 2, 1)
        line["cpu_baseline"] = {"value": len(times) / sum(times), "unit": "stacks/s", "cores": cores, "kind": "port",
                                "sample": "2 timed forwards of 1 stack (10x3x384x576) after 1 warm-up, oracle port of the "
                                          "reference's torch CPU path, %d torch threads" % cores}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DFF_BENCH_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--micro-batch", type=int, default=64)
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--e2e-micro-batch", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        import __graft_entry__ as g
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            g.build()
        run_ours(args)


if __name__ == "__main__":
    main()
