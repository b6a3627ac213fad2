"""One training step of the reference's recipe as a reusable driver (train_codes/train_code_Defocus.py:158-168):

    mid_out, pred1, pred2, pred3 = model(train_input, focus_dists)        # train-mode BatchNorm
    optimizer.zero_grad()
    Total = 0.5*MSE(pred1[mask], gt[mask]) + 0.7*MSE(pred2..) + 1.0*MSE(pred3..) + 0.3*MSE(mid_out..)      (:160-165)
    Total.backward(); optimizer.step()                                    # Adam(lr, betas=(0.9, 0.99)), :67

What it replaces around the model call: the four boolean-mask gathers (`nonzero` + `index` + their backward scatter, a radix sort
inside `index_put_` backward) and the ~190-tensor `torch.optim.Adam` by ONE masked-MSE kernel over the four heads (loss and
d loss / d pred in one pass, `dff_masked_mse`) and ONE Adam kernel over the flat parameter / gradient buffers (`dff_adam_flat`,
bit-compatible with `torch.optim.Adam`'s CUDA default, the foreach implementation); across ranks, `nn.DataParallel`'s per-step parameter
broadcast and gradient reduce by one all-reduce of the flat gradient bucket (distributed.py), launched on a side stream so that
it overlaps the optimizer bookkeeping of the rank.  The model call itself is the drop-in module (train.py behind it).

Loss normalisation across ranks follows the reference, which takes each MSE mean over the valid pixels of the *gathered global*
batch: every rank scales by its own valid-pixel count n_r and the bucket is divided by sum(n_r) after the all-reduce.
"""
import ctypes

import torch
import torch.distributed as dist

from . import distributed as D
from . import runtime as rt

_P = ctypes.c_void_p


def _declare(lib):
    if getattr(lib, "_dff_step_declared", False):
        return
    c = ctypes
    vp, i, i64, f, d = c.c_void_p, c.c_int, c.c_int64, c.c_float, c.c_double
    lib.dff_masked_mse.restype = i
    lib.dff_masked_mse.argtypes = [c.POINTER(vp), vp, vp, i64, c.POINTER(f), c.POINTER(vp), vp, vp, i, vp]
    lib.dff_adam_flat.restype = i
    lib.dff_adam_flat.argtypes = [vp, vp, vp, vp, i64, d, d, d, d, i, vp, i, vp]
    lib._dff_step_declared = True


class MaskedMSE(torch.autograd.Function):
    """sum_k w_k * mean_{mask}((pred_k - gt)^2) over the four heads, forward and backward in one kernel launch."""

    @staticmethod
    def forward(ctx, gt, mask, weights, p0, p1, p2, p3):
        l = rt.lib()
        _declare(l)
        dev = p0.device
        preds = [p.contiguous() for p in (p0, p1, p2, p3)]
        n = preds[0].numel()
        grads = [torch.empty_like(p) for p in preds]
        stats = torch.zeros(8, dtype=torch.float32, device=dev)          # [count, loss, per-head sums...]
        scratch = torch.empty(8 * 1024, dtype=torch.float64, device=dev)
        pp = (_P * 4)(*[p.data_ptr() for p in preds])
        gp = (_P * 4)(*[g.data_ptr() for g in grads])
        w = (ctypes.c_float * 4)(*[float(x) for x in weights])
        m8 = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8).contiguous()
        rt.check(l.dff_masked_mse(pp, _P(gt.contiguous().data_ptr()), _P(m8.data_ptr()), n, w, gp, _P(stats.data_ptr()),
                                  _P(scratch.data_ptr()), dev.index, _P(torch.cuda.current_stream(dev).cuda_stream)))
        ctx.save_for_backward(*grads)
        ctx.count = stats[0:1]
        return stats[1].clone()

    @staticmethod
    def backward(ctx, g):
        grads = ctx.saved_tensors
        return (None, None, None) + tuple(gr * g for gr in grads)


def masked_mse_loss(outs, gt, mask, weights):
    """Total loss of the reference recipe for outs = (mid_out, pred1, pred2, pred3) and weights in the same order."""
    return MaskedMSE.apply(gt, mask, tuple(weights), *outs)


class TrainStep:
    """Flat parameter / gradient / Adam-state buffers + the step of the reference's training loop."""

    def __init__(self, model, lr, betas=(0.9, 0.99), eps=1e-8, weights=(0.3, 0.5, 0.7, 1.0), group=None, use_graph=True):
        self.model, self.lr, self.betas, self.eps, self.weights, self.group = model, lr, betas, eps, weights, group
        # forward + loss + backward of a fixed-shape step are ~1500 kernel launches behind ~3000 Python-level operator calls: after two
        # eager steps they are captured into ONE CUDA graph (static input buffers) and replayed, which takes the host out of the loop.
        # The gradient all-reduce and the Adam kernel (whose bias corrections change every step) stay outside the graph.
        self.use_graph = use_graph and not isinstance(model, torch.nn.DataParallel)
        self._graph, self._static, self._eager_steps = None, None, 0
        net = model.module if isinstance(model, torch.nn.DataParallel) else model
        self.skip = D.unused_parameter_names(net)
        self.bucket = D.GradBucket(net, skip=self.skip)
        params = self.bucket.params
        dev = params[0].device
        # parameters become views of one flat fp32 buffer (state_dict / checkpoints are unaffected: same tensors, same values)
        self.flat_p = torch.empty(self.bucket.numel, dtype=torch.float32, device=dev)
        o = 0
        for p in params:
            n = p.numel()
            self.flat_p[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[o:o + n].view_as(p)
            o += n
        for p in params:          # weight / BatchNorm gradients are written straight into their bucket slots (train.py)
            p._dff_grad_sink = True
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.t = 0
        self.allreduce_bytes = 4 * (self.bucket.numel + 1)
        self.side = torch.cuda.Stream(device=dev)
        import os
        from . import train as _tr
        self.wgrad_side = _tr.WgradSideStream(dev) if os.environ.get("DFF_B200_WGRAD_STREAM", "1") != "0" else None
        self._ar_events = []
        rt.packed_cache(net.DFF_net if hasattr(net, "DFF_net") else net).invalidate()

    def _adam(self):
        l = rt.lib()
        _declare(l)
        self.t += 1
        dev = self.flat_p.device
        b1, b2 = self.betas
        rt.check(l.dff_adam_flat(_P(self.flat_p.data_ptr()), _P(self.bucket.flat.data_ptr()), _P(self.exp_avg.data_ptr()),
                                 _P(self.exp_avg_sq.data_ptr()), self.bucket.numel, self.lr, b1, b2, self.eps, self.t, None,
                                 dev.index, _P(torch.cuda.current_stream(dev).cuda_stream)))
        # in-place writes through raw pointers: make them visible to version-keyed caches
        for p in self.bucket.params:
            torch.autograd.graph.increment_version(p)

    def allreduce_ms(self):
        """Mean device time of the gradient all-reduce (scale, all-reduce, normalise) over the steps timed so far."""
        if not self._ar_events:
            return 0.0
        self._ar_events[-1][1].synchronize()
        return sum(a.elapsed_time(b) for a, b in self._ar_events) / len(self._ar_events)

    def _fwd_bwd(self, FS, fd, gt, mask):
        outs = self.model(FS, fd)
        self.bucket.zero()
        loss = masked_mse_loss(outs, gt, mask, self.weights)
        if self.wgrad_side is not None:
            with self.wgrad_side:          # weight gradients on a side stream, joined before anything reads the bucket
                loss.backward()
        else:
            loss.backward()
        return loss.detach()

    def _fwd_bwd_graphed(self, FS, fd, gt, mask):
        sig = tuple((tuple(t.shape), t.dtype, tuple(t.stride())) for t in (FS, fd, gt, mask))
        if self._graph is not None and self._static["sig"] == sig:
            st = self._static
            for dst, src in zip(st["in"], (FS, fd, gt, mask)):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src)
            self._graph.replay()
            return st["loss"]
        if self._eager_steps < 2 or self._graph is False:
            self._eager_steps += 1
            return self._fwd_bwd(FS, fd, gt, mask)
        try:
            self.bucket.rebind()
            static_in = [t.clone() for t in (FS, fd, gt, mask)]
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph):
                loss = self._fwd_bwd(*static_in)
            self._graph, self._static = graph, {"sig": sig, "in": static_in, "loss": loss}
            return loss        # (capture does not execute: the caller gets this step's numbers from the replay below)
        except Exception as ex:   # capture is an optimisation: report once, keep training eagerly
            import warnings
            warnings.warn("dff_b200: CUDA-graph capture of the training step failed (%s: %s); running eagerly" % (type(ex).__name__, ex))
            self._graph = False
            torch.cuda.synchronize()
            return self._fwd_bwd(FS, fd, gt, mask)

    def step(self, FS, fd, gt, mask, time_allreduce=False):
        info = {}
        if self.use_graph:
            captured_before = self._graph not in (None, False)
            loss = self._fwd_bwd_graphed(FS, fd, gt, mask)
            if not captured_before and self._graph not in (None, False):
                self._graph.replay()      # the step that captured the graph has not run yet
        else:
            loss = self._fwd_bwd(FS, fd, gt, mask)
        world = dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1
        if world > 1:
            n_r = mask.sum().float()
            cur = torch.cuda.current_stream()
            if time_allreduce:
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                if time_allreduce:
                    a0.record()
                self.bucket.flat[:self.bucket.numel].mul_(n_r)
                self.bucket.flat[self.bucket.numel] = n_r
                dist.all_reduce(self.bucket.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.bucket.flat[:self.bucket.numel].div_(self.bucket.flat[self.bucket.numel].clamp_min(1e-30))
                if time_allreduce:
                    a1.record()
            cur.wait_stream(self.side)
            if time_allreduce:
                self._ar_events.append((a0, a1))     # read after the timed loop (no host sync inside a step)
        self._adam()
        info["loss"] = loss
        info["graphed"] = self._graph not in (None, False)
        return info
