"""The GPU incumbent: the reference's dataflow in torch eager on one B200 — cuDNN / ATen
kernels, `cudnn.benchmark=True` as train_code_Defocus.py:76 sets it — in the three configurations a user of the reference could run:
   tf32       fp32 tensors, torch defaults (cuDNN convolutions silently use TF32)      <- what `python test.py` does today
   fp32       strict fp32 (allow_tf32 off)                                             <- the numerically comparable run
   bf16_cl3d  bf16 weights/activations, channels_last_3d                               <- the fastest the stock stack offers
The dataflow is `oracle/dff_oracle.py` (bit-identical to the reference module, tests/test_oracle_golden.py) fed CUDA tensors; this
is a measurement tool, not part of the product path.  Prints one JSON document (and writes it with --out).

    python tools/incumbent.py [--batch 8] [--steps 5] [--layers] [--out profiles/r2_gpu_incumbent.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from dffinthewild_b200 import synth  # noqa: E402
from dffinthewild_b200.Depth_Estimation_Network import Network  # noqa: E402
from oracle import dff_oracle as O  # noqa: E402

S, H, W = 10, 384, 576


def timed(fn, steps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--layers", action="store_true", help="also time the cuDNN kernels of the slowest layers in isolation")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    sd = synth.synthetic_state(Network().state_dict(), seed=1)
    FS = synth.focal_stack(a.batch, S, H, W, seed=0, valid_hw=(383, 552)).to(dev)
    fd = synth.focus_dists(a.batch, S, H, W, "ddff").to(dev)
    res = {"what": "reference dataflow in torch eager (cuDNN %s, torch %s), cudnn.benchmark=True, %d DDFF stacks (10x3x384x576) per call, "
                   "CUDA-event timed, inputs resident" % (torch.backends.cudnn.version(), torch.__version__, a.batch),
           "gpu": torch.cuda.get_device_name(0), "batch": a.batch, "steps": a.steps, "modes": {}}
    for mode in ("tf32", "fp32", "bf16_cl3d"):
        tf32 = mode == "tf32"
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        dt = torch.bfloat16 if mode == "bf16_cl3d" else torch.float32
        sdm = {k: (v.to(dev, dt) if v.is_floating_point() else v.to(dev)) for k, v in sd.items()}
        x, f = FS.to(dt), fd.to(dt)
        if mode == "bf16_cl3d":
            x = x.contiguous(memory_format=torch.channels_last_3d)
            sdm = {k: (v.contiguous(memory_format=torch.channels_last_3d) if v.dim() == 5 else v) for k, v in sdm.items()}
        try:
            with torch.no_grad():
                ms = timed(lambda: O.dff_forward(sdm, x, f), a.steps)
            res["modes"][mode] = {"ms_per_call": ms, "stacks_per_s": a.batch / (ms / 1e3),
                                  "tflops": 93563.0 * S * H * W * a.batch / (ms / 1e3) / 1e12}
        except Exception as ex:   # (e.g. an op without a bf16 kernel) — reported, not hidden
            res["modes"][mode] = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}
        torch.cuda.empty_cache()
    if a.layers:
        # the cuDNN kernels of the layers that dominate our own step (profiles/r1_by_layer_bf16.txt), isolated: conv only, no BN/ReLU
        B = a.batch
        layers = [("dres4.conv0 (16->8 3x3x3 @1/1)", 16, 8, 3, 1, 1, False, H, W),
                  ("FM_measure.0 (3->8 1x9x9 dil2 @1/1)", 3, 8, 9, 1, 2, False, H, W),
                  ("dres3.conv0 (32->16 3x3x3 @1/2)", 32, 16, 3, 1, 1, False, H // 2, W // 2),
                  ("dres2.conv0 (64->32 3x3x3 @1/4)", 64, 32, 3, 1, 1, False, H // 4, W // 4),
                  ("dres4.conv6 (deconv 16->8 -> 1/1)", 16, 8, 3, 2, 1, True, H // 2, W // 2),
                  ("dres0.2 (64->64 3x3x3 @1/8)", 64, 64, 3, 1, 1, False, H // 8, W // 8)]
        out = {}
        for name, ci, co, k, st, dil, tr, h, w in layers:
            row = {}
            for mode in ("tf32", "bf16_cl3d"):
                torch.backends.cudnn.allow_tf32 = True
                dt = torch.bfloat16 if mode == "bf16_cl3d" else torch.float32
                x = torch.randn(B, ci, S, h, w, device=dev, dtype=dt)
                if k == 9:
                    wt = torch.randn(co, ci, 1, 9, 9, device=dev, dtype=dt)
                    fn = lambda: F.conv3d(x, wt, None, 1, (0, 8, 8), (1, 2, 2))
                elif tr:
                    wt = torch.randn(ci, co, 3, 3, 3, device=dev, dtype=dt)
                    fn = lambda: F.conv_transpose3d(x, wt, None, stride=(1, 2, 2), padding=1, output_padding=(0, 1, 1))
                else:
                    wt = torch.randn(co, ci, 3, 3, 3, device=dev, dtype=dt)
                    fn = lambda: F.conv3d(x, wt, None, (1, st, st), 1)
                if mode == "bf16_cl3d":
                    x = x.contiguous(memory_format=torch.channels_last_3d)
                    wt = wt.contiguous(memory_format=torch.channels_last_3d)
                try:
                    with torch.no_grad():
                        row[mode + "_us"] = 1e3 * timed(fn, a.steps)
                except Exception as ex:
                    row[mode + "_us"] = "%s" % type(ex).__name__
                del x, wt
                torch.cuda.empty_cache()
            out[name] = row
        res["layers_cudnn_conv_only"] = out
    print(json.dumps(res))
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
