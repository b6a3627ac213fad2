"""Train-mode forward/backward of `DFF_net` (reference train_codes/Depth_Estimation_Network.py:77-137 under `model.train()`,
backward invoked by `Total.backward()` at train_codes/train_code_Defocus.py:167).

PyTorch's autograd is the *tape* (which operator produced which tensor, gradient accumulation where a tensor has several
consumers); every operator's forward and backward computation is a call into the C-ABI library (`include/dff_b200.h`, section
"train-mode building blocks").  Activations are channels-last `(B,S,H,W,C)` fp32 tensors (parity mode); parameter gradients
land in the module's own `nn.Parameter.grad`, so `torch.optim.Adam(model.parameters())` and the reference's training scripts
work unchanged.  BatchNorm runs in batch-statistics mode and updates `running_mean/var` (momentum 0.1, unbiased variance) and
`num_batches_tracked` exactly like `nn.BatchNorm3d`.

Operators (autograd.Function):
  ConvFn      conv / transposed conv (two-source = torch.cat-free)   fwd dff_conv3d, bwd dff_conv3d_dgrad + dff_conv3d_wgrad
  BnActFn     [relu](BN_batch(x) + res_pre) + res_post               fwd dff_bn_train_forward, bwd dff_bn_train_backward
  PoolFn      (1,k,k) max / average pooling                          dff_pool3d / dff_pool3d_backward
  AddFn       skip add                                                dff_add (backward is the identity on both inputs)
  DepthHeadFn upsample + softplus-normalise + expected focus distance dff_depth_head / dff_depth_head_backward
"""
import ctypes

import torch

from . import runtime as rt

_P = ctypes.c_void_p


def _p(t):
    return _P(t.data_ptr()) if t is not None else _P(0)


def _declare_train(lib):
    if getattr(lib, "_dff_train_declared", False):
        return
    c = ctypes
    vp, i, i64, sz, f = c.c_void_p, c.c_int, c.c_int64, c.c_size_t, c.c_float
    sig = {
        "dff_conv3d_dgrad_scratch_bytes": (sz, [i, i, i, i, i]),
        "dff_conv3d_dgrad": (i, [vp, i, i, i, i, i, vp, i, i, i, i, i, i, i, i, i, i, vp, i, vp, i, vp]),
        "dff_conv3d_wgrad": (i, [vp, i, vp, i, i, i, i, i, vp, i, i, i, i, i, i, i, i, i, vp, i, i, vp]),
        "dff_conv3d_wgrad_acc": (i, [vp, i, vp, i, i, i, i, i, vp, i, i, i, i, i, i, i, i, i, vp, i, i, vp]),
        "dff_bn_scratch_bytes": (sz, [i]),
        "dff_bn_train_forward": (i, [vp, i64, i, i, vp, vp, vp, vp, f, f, vp, vp, i, vp, vp, vp, vp, vp, i, vp]),
        "dff_bn_train_backward": (i, [vp, vp, vp, vp, vp, vp, i64, i, i, vp, vp, vp, vp, vp, i, vp]),
        "dff_bn_eval_forward": (i, [vp, i64, i, i, vp, vp, vp, vp, f, vp, vp, i, vp, vp, vp, vp, i, vp]),
        "dff_bn_eval_backward": (i, [vp, vp, vp, vp, vp, vp, i64, i, i, vp, vp, vp, vp, vp, i, vp]),
        "dff_add": (i, [vp, vp, i64, i, vp, i, vp]),
        "dff_pool3d": (i, [vp, i, i, i, i, i, i, i, vp, i, vp]),
        "dff_pool3d_backward": (i, [vp, vp, i, i, i, i, i, i, i, vp, i, vp]),
        "dff_depth_head_backward": (i, [vp, i, i, vp, c.POINTER(i64), i, i, i, i, vp, vp, i, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    lib._dff_train_declared = True


def _lib():
    l = rt.lib()
    _declare_train(l)
    return l


def _st(dev):
    return _P(torch.cuda.current_stream(dev).cuda_stream)


def _elem(t):
    return rt.BF16 if t.dtype == torch.bfloat16 else rt.FP32


OUT_F32 = 16   # DFF_OUT_F32

_bn_scratch_cache = {}
_pending_nbt = []     # num_batches_tracked buffers of the BatchNorm layers of the running forward: bumped with ONE foreach add


def _bn_scratch(l, C, dev):
    """Scratch of the BatchNorm reductions (per-block partial sums): one buffer per (device, stream, C) serves every call — calls
    on a stream are ordered, and a buffer created outside a CUDA-graph capture may be used inside one."""
    st = torch.cuda.current_stream(dev)
    key = (dev.index, st.cuda_stream, C)
    buf = _bn_scratch_cache.get(key)
    if buf is None:
        buf = torch.empty(l.dff_bn_scratch_bytes(C), dtype=torch.uint8, device=dev)
        if torch.cuda.is_current_stream_capturing():
            return buf      # private-pool memory of the capture: not reusable outside it
        _bn_scratch_cache[key] = buf
    return buf


class WgradSideStream:
    """Context of a training step that joins before it reads the gradients (train_step.TrainStep): inside it the weight-gradient
    kernels — which nothing in the backward pass consumes — go to a side stream and fill the gaps of the data-gradient chain
    (hundreds of 5-50 us kernels on the main stream).  Operands are kept alive until `join()`; without this context every kernel
    stays on the caller's stream, so `loss.backward(); optimizer.step()` of the reference scripts needs no extra synchronisation."""
    active = None

    def __init__(self, dev):
        self.dev, self.side, self.keep = dev, torch.cuda.Stream(device=dev), []

    def __enter__(self):
        WgradSideStream.active = self
        return self

    def __exit__(self, *a):
        WgradSideStream.active = None
        self.join()

    def fork(self, *tensors):
        self.side.wait_stream(torch.cuda.current_stream(self.dev))
        self.keep.append(tensors)
        return _P(self.side.cuda_stream)

    def join(self):
        if self.keep:
            torch.cuda.current_stream(self.dev).wait_stream(self.side)
            del self.keep[:]


def flush_batch_counters():
    """nn.BatchNorm3d.num_batches_tracked += 1 for every BatchNorm that ran since the last flush (one launch instead of 58)."""
    if _pending_nbt:
        torch._foreach_add_(_pending_nbt, 1)
        del _pending_nbt[:]


def _conv_call(l, x0, x1, w, Cout, stride, dil, transposed, out, elem, tc):
    """One dff_conv3d call on channels-last tensors (no epilogue operands)."""
    dev = x0.device
    B, S, IH, IW, C0 = x0.shape
    C1 = x1.shape[-1] if x1 is not None else 0
    kd, kh, kw = w.shape[2:]
    scratch = torch.empty(l.dff_conv3d_scratch_bytes(C0 + C1, Cout, kd, kh, kw), dtype=torch.uint8, device=dev)
    rt.check(l.dff_conv3d(_p(x0), C0, _p(x1), C1, B, S, IH, IW, _p(w), Cout, kd, kh, kw, stride, dil, 1 if transposed else 0, None, None,
                          None, None, 0, _p(out), elem, tc, _p(scratch), dev.index, _st(dev)))


class ConvFn(torch.autograd.Function):
    """Raw convolution (no bias): x0 (B,S,IH,IW,C0) [+ x1 = virtual channel concat], weight in the reference layout.
    fp32 tensors: FFMA kernels.  bf16 tensors: tcgen05 kernels for the forward AND the data gradient (a data gradient is a
    convolution with the adjoint tap table: flipped taps for stride 1, a transposed convolution for stride 2 and vice versa);
    the weight gradient is always accumulated in fp32."""

    @staticmethod
    def forward(ctx, x0, x1, weight, stride, dil, transposed, cin_pad):
        l = _lib()
        dev = x0.device
        bf16 = x0.dtype == torch.bfloat16
        B, S, IH, IW, C0 = x0.shape
        w = weight
        if cin_pad:  # first layer: 3 image channels stored as 4 (fp32) / 8 (bf16); zero weights for the padding channels
            w = torch.cat([weight, weight.new_zeros(weight.shape[0], cin_pad, *weight.shape[2:])], 1)
        w = w.detach().float().contiguous()
        Cout = w.shape[1] if transposed else w.shape[0]
        OH, OW = (IH * 2, IW * 2) if transposed else (IH // stride, IW // stride)
        cost = Cout % 4 != 0   # C -> 1 cost volumes stay fp32 (they feed the depth head)
        out = torch.empty((B, S, OH, OW, Cout), dtype=torch.float32 if (cost or not bf16) else torch.bfloat16, device=dev)
        elem = (rt.BF16 | (OUT_F32 if cost else 0)) if bf16 else rt.FP32
        _conv_call(l, x0, x1, w, Cout, 2 if transposed else stride, dil, transposed, out, elem, 1 if bf16 else 0)
        ctx.save_for_backward(x0, x1, w)
        ctx.cfg = (stride, dil, transposed, cin_pad, Cout, OH, OW)
        # gradient sink (train_step.TrainStep): the weight gradient is accumulated straight into the parameter's slot of the flat,
        # once-per-step-zeroed bucket — no temporary, no AccumulateGrad add
        ctx.sink = weight if getattr(weight, "_dff_grad_sink", False) else None
        return out

    @staticmethod
    def backward(ctx, dy):
        l = _lib()
        x0, x1, w = ctx.saved_tensors
        stride, dil, transposed, cin_pad, Cout, OH, OW = ctx.cfg
        dev = dy.device
        bf16 = x0.dtype == torch.bfloat16
        elem = rt.BF16 if bf16 else rt.FP32
        dy = dy.contiguous()
        B, S, IH, IW, C0 = x0.shape
        C1 = x1.shape[-1] if x1 is not None else 0
        Cin = C0 + C1
        kd, kh, kw = w.shape[2:]
        st = stride if not transposed else 2
        CoS = dy.shape[-1]
        if CoS % 4:  # single-channel cost volumes (fp32): pad dy to 4 (fp32) / 8 (bf16) stored channels with the library's layout kernel
            cp = 8 if bf16 else 4
            dyp = torch.empty(dy.shape[:-1] + (cp,), dtype=x0.dtype, device=dev)
            rt.check(l.dff_to_channels_last(_p(dy.float()), B, CoS, S, OH, OW, _p(dyp), cp, elem, dev.index, _st(dev)))
            dy, CoS = dyp, cp       # dy (B,S,OH,OW,1) == (B,1,S,OH,OW) in memory
        needs = ctx.needs_input_grad
        dx0 = dx1 = None
        for idx, (x, ci0) in enumerate(((x0, 0), (x1, C0))):
            if x is None or not needs[idx]:
                continue
            nci = x.shape[-1]
            dx = torch.empty_like(x)
            # fp32: FFMA kernels with adjoint tap tables.  bf16: the library runs the adjoint convolution on the tcgen05 kernels (flipped
            # taps for stride 1, a transposed convolution for stride 2 and vice versa), re-laying the weight out with one small kernel.
            scratch = torch.empty(l.dff_conv3d_dgrad_scratch_bytes(Cin, max(Cout, CoS), kd, kh, kw), dtype=torch.uint8, device=dev)
            rt.check(l.dff_conv3d_dgrad(_p(dy), CoS, B, S, OH, OW, _p(w), Cin, Cout, kd, kh, kw, st, dil, 1 if transposed else 0,
                                        ci0, nci, _p(dx), elem, _p(scratch), dev.index, _st(dev)))
            if idx == 0:
                dx0 = dx
            else:
                dx1 = dx
        dw = None
        sink = ctx.sink
        if needs[2] and sink is not None and sink.grad is not None and sink.grad.is_contiguous() and sink.grad.dtype == torch.float32:
            # (stored channels beyond the layer's Cin — the first layer's padding — are skipped by the kernel)
            cin_true = sink.shape[0] if transposed else sink.shape[1]
            ws = WgradSideStream.active
            stream = ws.fork(x0, x1, dy) if (ws is not None and ws.dev == dev) else _st(dev)
            rt.check(l.dff_conv3d_wgrad_acc(_p(x0), C0, _p(x1), C1, B, S, IH, IW, _p(dy), CoS, cin_true, Cout, kd, kh, kw, st, dil,
                                            1 if transposed else 0, _p(sink.grad), elem, dev.index, stream))
        elif needs[2]:
            dw = torch.empty_like(w)
            rt.check(l.dff_conv3d_wgrad(_p(x0), C0, _p(x1), C1, B, S, IH, IW, _p(dy), CoS, Cin, Cout, kd, kh, kw, st, dil,
                                        1 if transposed else 0, _p(dw), elem, dev.index, _st(dev)))
            if cin_pad:
                dw = dw[:, :Cin - cin_pad].contiguous()
        return dx0, dx1, dw, None, None, None, None


class ToChannelsLastFn(torch.autograd.Function):
    """FS (B,3,S,H,W) fp32 -> channels-last (B,S,H,W,Cp): only on the tape when the caller wants d loss / d FS."""

    @staticmethod
    def forward(ctx, FS, Cp, bf16):
        ctx.C = FS.shape[1]
        return rt.to_channels_last(FS, Cp, bf16)

    @staticmethod
    def backward(ctx, dx):
        return rt.from_channels_last(dx.contiguous(), ctx.C), None, None


class BnActFn(torch.autograd.Function):
    """out = [relu](BN_batchstats(x) + res_pre) + res_post.  bn = the module's nn.BatchNorm3d (None: no normalisation)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, res_pre, res_post, relu, bn):
        l = _lib()
        dev = x.device
        C = x.shape[-1]
        npix = x.numel() // C
        out = torch.empty_like(x)
        mean = invstd = None
        batch_stats = bn is None or bn.training or bn.running_mean is None   # (torch: no running statistics => batch statistics)
        if gamma is not None:
            mean = torch.empty(C, dtype=torch.float32, device=dev)
            invstd = torch.empty(C, dtype=torch.float32, device=dev)
            ss = torch.empty(2 * C, dtype=torch.float32, device=dev)
            if batch_stats:
                scratch = _bn_scratch(l, C, dev)
                track = bn is not None and bn.track_running_stats and bn.running_mean is not None
                momentum = 0.1
                if bn is not None and track:
                    # nn.BatchNorm3d: momentum=None means a cumulative moving average over num_batches_tracked
                    momentum = bn.momentum if bn.momentum is not None else 1.0 / float(int(bn.num_batches_tracked) + 1)
                rt.check(l.dff_bn_train_forward(_p(x), npix, C, _elem(x), _p(gamma), _p(beta), _p(bn.running_mean) if track else None,
                                                _p(bn.running_var) if track else None, momentum,
                                                bn.eps if bn is not None else 1e-5, _p(res_pre), _p(res_post), 1 if relu else 0,
                                                _p(out), _p(mean), _p(invstd), _p(ss), _p(scratch), dev.index, _st(dev)))
                if track:
                    _pending_nbt.append(bn.num_batches_tracked)
                    # the library wrote the running statistics through raw pointers: make the update visible to version-keyed caches
                    torch.autograd.graph.increment_version(bn.running_mean)
                    torch.autograd.graph.increment_version(bn.running_var)
            else:   # eval-mode BatchNorm inside a taped forward: running statistics, no update
                rt.check(l.dff_bn_eval_forward(_p(x), npix, C, _elem(x), _p(gamma), _p(beta), _p(bn.running_mean), _p(bn.running_var),
                                               bn.eps, _p(res_pre), _p(res_post), 1 if relu else 0, _p(out), _p(mean), _p(invstd),
                                               _p(ss), dev.index, _st(dev)))
        else:
            rt.check(l.dff_bn_train_forward(_p(x), npix, C, _elem(x), None, None, None, None, 0.0, 0.0, _p(res_pre), _p(res_post),
                                            1 if relu else 0, _p(out), None, None, None, None, dev.index, _st(dev)))
        # ReLU mask source: the output before res_post.  Without res_post that is `out` itself; with it, relu(x) > 0 <=> x > 0
        # (no-BN layers, reference :399-402) so the raw input serves as the mask.
        ctx.relu, ctx.has_post, ctx.batch_stats = relu, res_post is not None, batch_stats
        sink = gamma is not None and getattr(gamma, "_dff_grad_sink", False) and getattr(beta, "_dff_grad_sink", False)
        ctx.sinks = (gamma, beta) if sink else (None, None)
        ctx.save_for_backward(x, gamma, mean, invstd, out if (relu and res_post is None) else None)
        return out

    @staticmethod
    def backward(ctx, dy):
        l = _lib()
        x, gamma, mean, invstd, y = ctx.saved_tensors
        dev = dy.device
        dy = dy.to(x.dtype).contiguous()
        C = x.shape[-1]
        npix = x.numel() // C
        mask = None
        if ctx.relu:
            mask = y if y is not None else x
            if y is None and gamma is not None:
                raise rt.DffError("dff_b200: relu + res_post after BatchNorm is not a pattern of this network")
        needs = ctx.needs_input_grad
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if needs[3] else None
        dgamma = dbeta = None
        sunk = False
        if gamma is not None:
            gs, bs_ = ctx.sinks
            if gs is not None and gs.grad is not None and bs_.grad is not None and gs.grad.is_contiguous() and bs_.grad.is_contiguous():
                dgamma, dbeta, sunk = gs.grad, bs_.grad, True      # every BatchNorm runs once per step: plain overwrite of its slot
            else:
                dgamma = torch.empty(C, dtype=torch.float32, device=dev)
                dbeta = torch.empty(C, dtype=torch.float32, device=dev)
            scratch = _bn_scratch(l, C, dev)
            fn = l.dff_bn_train_backward if ctx.batch_stats else l.dff_bn_eval_backward
            rt.check(fn(_p(dy), _p(mask), _p(x), _p(mean), _p(invstd), _p(gamma), npix, C, _elem(x), _p(dx), _p(dres), _p(dgamma), _p(dbeta),
                        _p(scratch), dev.index, _st(dev)))
        else:
            rt.check(l.dff_bn_train_backward(_p(dy), _p(mask), None, None, None, None, npix, C, _elem(x), _p(dx), _p(dres), None, None,
                                             None, dev.index, _st(dev)))
        if sunk:
            dgamma = dbeta = None
        return dx, dgamma, dbeta, dres, (dy if ctx.has_post else None), None, None


class PoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, is_max):
        l = _lib()
        B, S, H, W, C = x.shape
        out = torch.empty((B, S, H // k, W // k, C), dtype=x.dtype, device=x.device)
        rt.check(l.dff_pool3d(_p(x), B * S, H, W, C, k, 1 if is_max else 0, _elem(x), _p(out), x.device.index, _st(x.device)))
        ctx.save_for_backward(x)
        ctx.cfg = (k, is_max)
        return out

    @staticmethod
    def backward(ctx, dy):
        l = _lib()
        (x,) = ctx.saved_tensors
        k, is_max = ctx.cfg
        B, S, H, W, C = x.shape
        dx = torch.empty_like(x)
        rt.check(l.dff_pool3d_backward(_p(x), _p(dy.contiguous()), B * S, H, W, C, k, 1 if is_max else 0, _elem(x), _p(dx),
                                       x.device.index, _st(x.device)))
        return dx, None, None


class AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        l = _lib()
        out = torch.empty_like(a)
        rt.check(l.dff_add(_p(a), _p(b), a.numel(), _elem(a), _p(out), a.device.index, _st(a.device)))
        return out

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


class DepthHeadFn(torch.autograd.Function):
    """cost (B,S,h,w) fp32 (the C=1 channels-last cost volume), fd strided view (B,S,H,W) -> depth (B,H,W)."""

    @staticmethod
    def forward(ctx, cost, fd, H, W):
        l = _lib()
        B, S, h, w = cost.shape
        out = torch.empty((B, H, W), dtype=torch.float32, device=cost.device)
        strides = (ctypes.c_int64 * 4)(*fd.stride())
        rt.check(l.dff_depth_head(_p(cost), h, w, _p(fd), strides, B, S, H, W, _p(out), cost.device.index, _st(cost.device)))
        ctx.save_for_backward(cost, fd)
        ctx.hw = (H, W)
        return out

    @staticmethod
    def backward(ctx, dd):
        l = _lib()
        cost, fd = ctx.saved_tensors
        H, W = ctx.hw
        B, S, h, w = cost.shape
        dcost = torch.empty_like(cost)
        strides = (ctypes.c_int64 * 4)(*fd.stride())
        rt.check(l.dff_depth_head_backward(_p(cost), h, w, _p(fd), strides, B, S, H, W, _p(dd.contiguous()), _p(dcost),
                                           cost.device.index, _st(cost.device)))
        return dcost, None, None, None


# ---------------------------------------------------------------------------------------------------------------
# the network (same dataflow as runtime/net.cu's eval schedule; module attribute paths as in the reference)
# ---------------------------------------------------------------------------------------------------------------
def _conv(x, conv, x1=None, cin_pad=0):
    transposed = isinstance(conv, torch.nn.ConvTranspose3d)
    stride = 2 if transposed else conv.stride[1]
    return ConvFn.apply(x, x1, conv.weight, stride, conv.dilation[1], transposed, cin_pad)


def _cbn(x, seq, relu=False, res_pre=None, res_post=None, x1=None, cin_pad=0):
    """`convbn_3d` (reference :352-355): seq[0] conv, seq[1] BatchNorm3d."""
    bn = seq[1]
    return BnActFn.apply(_conv(x, seq[0], x1, cin_pad), bn.weight, bn.bias, res_pre, res_post, relu, bn)


def _act(x, relu=True, res_post=None):
    return BnActFn.apply(x, None, None, None, res_post, relu, None)


def _srd(m, x):
    """Feature_Extraction / SRD (reference :394-407)."""
    fm = m.Focus_Measure.conv
    t = _cbn(x, fm[0], relu=True)
    f = _cbn(t, fm[2], relu=True, res_pre=x)
    a = _act(_conv(f, m.N_ch_attention[0]))
    return _act(_conv(a, m.N_ch_attention[2]), res_post=f)


def _efd(m, x):
    """res_stride_conv_3d / EFD (reference :383-392)."""
    a = _cbn(x, m.stride_conv)
    mp = PoolFn.apply(x, 2, True)
    return _cbn(mp, m.max_pooling[1], relu=True, res_pre=a)


def _pyramid(m, x):
    """hourglassup (reference :247-273)."""
    def tower(t, s0, s1):
        r = _cbn(_cbn(t, s0[0], relu=True), s0[2], relu=True)
        return _cbn(_cbn(r, s1[0], relu=True), s1[2], res_pre=r)

    x8 = tower(PoolFn.apply(x, 2, False), m.dres8_0, m.dres8_1)
    x16 = tower(PoolFn.apply(x, 4, False), m.dres16_0, m.dres16_1)
    x32 = tower(PoolFn.apply(x, 8, False), m.dres32_0, m.dres32_1)
    c1 = _conv(x8, m.conv1)
    c1 = _cbn(c1, m.combine1[0], relu=True, x1=x16)
    c2 = _cbn(c1, m.conv2[0], relu=True)
    c3 = _conv(c2, m.conv3)
    c3 = _cbn(c3, m.combine2[0], relu=True, x1=x32)
    c4 = _cbn(c3, m.conv4[0], relu=True)
    c8 = _cbn(c4, m.conv8, relu=True, res_pre=_cbn(c2, m.redir2))
    return _cbn(c8, m.conv9, relu=True, res_pre=_cbn(x8, m.redir1))


def _hourglass(m, x, skip, presqu, postsqu):
    """hourglass (reference :302-321) -> (out, pre_1)."""
    pre_1 = _cbn(x, m.conv0[0], relu=True, x1=skip)
    out = _cbn(pre_1, m.conv1[0], relu=True)
    pre = _cbn(out, m.conv2, relu=True, res_pre=postsqu)
    out = _cbn(pre, m.conv3[0], relu=True)
    out = _cbn(out, m.conv4[0], relu=True)
    out = _cbn(out, m.conv5, relu=True, res_pre=presqu if presqu is not None else pre)
    out = _cbn(out, m.conv6)
    return out, pre_1


def dff_net_train_forward(net, FS, focus_dists):
    """Train-mode `DFF_net.forward` (reference :77-137): returns (mid_out, pred1, pred2, pred3), autograd-connected."""
    rt._require_cuda(FS, "FS")
    rt._require_cuda(focus_dists, "focus_dists")
    if FS.dim() != 5 or FS.shape[1] != 3:
        raise rt.DffError("dff_b200: FS must be (B,3,S,H,W), got %s" % (tuple(FS.shape),))
    if FS.dtype != torch.float32 or focus_dists.dtype != torch.float32:
        raise rt.DffError("dff_b200: FS and focus_dists must be float32")
    bf16 = getattr(net, "precision", "fp32") == "bf16"
    B, _, S, H, W = FS.shape
    if H % 32 or W % 32:
        raise rt.DffError("dff_b200: H and W must be multiples of 32 (pad with -1 like the reference dataloaders)")
    dev = FS.device
    rt._check_device(dev.index)
    del _pending_nbt[:]
    fd = focus_dists
    while fd.dim() < 4:
        fd = fd.unsqueeze(0)
    fd = fd.expand(B, S, H, W)
    # (B,S,H,W,4|8): image channels + zero padding
    x0 = ToChannelsLastFn.apply(FS, 8 if bf16 else 4, bf16) if FS.requires_grad else rt.to_channels_last(FS, 8 if bf16 else 4, bf16)
    fm = net.FM_measure.Focus_extraction
    v1 = _srd(fm[2], _cbn(x0, fm[0], relu=True, cin_pad=5 if bf16 else 1))
    v2 = _srd(net.FM_conv1[1], _efd(net.FM_conv1[0], v1))
    v3 = _srd(net.FM_conv2[1], _efd(net.FM_conv2[0], v2))
    vol = _pyramid(net.SPP_module, v3)

    def head(cost):                                              # (B,S,h,w,1) -> depth
        return DepthHeadFn.apply(cost.reshape(cost.shape[:4]), fd, H, W)

    mid = head(_conv(_cbn(vol, net.confidence[0], relu=True), net.confidence[2]))
    x = _cbn(_cbn(vol, net.dres0[0], relu=True), net.dres0[2], relu=True)
    x = _cbn(x, net.deconv_1)
    out, pre = _hourglass(net.dres2, x, v3, None, None)
    out_in = AddFn.apply(x, out)
    p1 = head(_conv(out_in, net.classif1[0]))
    o2 = _cbn(out_in, net.deconv_2)
    outb, pre2 = _hourglass(net.dres3, o2, v2, pre, out)
    out_in2 = AddFn.apply(o2, outb)
    p2 = head(_conv(out_in2, net.classif2[0]))
    o3 = _cbn(out_in2, net.deconv_3)
    outc, _ = _hourglass(net.dres4, o3, v1, pre2, outb)
    p3 = head(_conv(AddFn.apply(o3, outc), net.classif3[0]))
    flush_batch_counters()
    return mid, p1, p2, p3
