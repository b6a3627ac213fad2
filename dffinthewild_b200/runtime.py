"""ctypes binding of the C-ABI library (`include/dff_b200.h`) and the glue the drop-in modules use.

Torch is plumbing here: it owns device memory (inputs, outputs, packed weights, workspace) and the current
stream; every computation happens inside `libdff_b200.so`.  There is deliberately no fallback: if the library is
missing, the device is not sm_100, or a tensor lives on the CPU, the call raises.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DFF_B200_LIB") or os.path.join(_HERE, "libdff_b200.so")  # (override: instrumented debug builds)

FP32, BF16, TRAIN, NO_TC, NO_SLAB = 0, 1, 2, 4, 8
NET_DFF, NET_FLOW = 0, 1

_lib = None
_lock = threading.Lock()


class DffError(RuntimeError):
    pass


def _declare(lib):
    c = ctypes
    vp, i, i64, sz, fp = c.c_void_p, c.c_int, c.c_int64, c.c_size_t, c.c_void_p
    sig = {
        "dff_abi_version": (i, []),
        "dff_last_error": (c.c_char_p, []),
        "dff_check_device": (i, [i]),
        "dff_param_count": (i, [i]),
        "dff_param_name": (c.c_char_p, [i, i]),
        "dff_param_numel": (i64, [i, i]),
        "dff_param_offset": (i64, [i, i]),
        "dff_raw_numel": (i64, [i]),
        "dff_packed_bytes": (sz, [i]),
        "dff_pack_weights": (i, [i, fp, vp, i, vp]),
        "dff_workspace_bytes": (sz, [i, i, i, i, i]),
        "dff_forward": (i, [vp, fp, fp, c.POINTER(i64), i, i, i, i, c.POINTER(vp), c.POINTER(vp), vp, sz, i, i, vp]),
        "dff_forward_profiled": (i, [vp, fp, fp, c.POINTER(i64), i, i, i, i, c.POINTER(vp), vp, sz, i, i, vp, i, vp, vp, vp, vp,
                                     vp, c.POINTER(i)]),
        "dff_host_io_bytes": (sz, [i, i, i, i]),
        "dff_forward_host": (i, [vp, fp, fp, c.POINTER(i64), i, i, i, i, i, c.POINTER(vp), vp, vp, sz, i, i, vp]),
        "dff_forward_u8": (i, [vp, vp, i, i, fp, c.POINTER(i64), i, i, i, i, c.POINTER(vp), c.POINTER(vp), vp, sz, i, i, vp]),
        "dff_host_io_bytes_u8": (sz, [i, i, i, i, i, i, c.POINTER(i64)]),
        "dff_forward_host_u8": (i, [vp, vp, i, i, fp, c.POINTER(i64), i, i, i, i, i, c.POINTER(vp), vp, vp, sz, i, i, vp]),
        "dff_forward_host_u8_async": (i, [vp, vp, i, i, fp, c.POINTER(i64), i, i, i, i, i, c.POINTER(vp), vp, vp, sz, i, i, vp, i]),
        "dff_forward_host_wait": (i, [i, i]),
        "dff_stage_u8": (i, [vp, i, i, i, i, i, i, fp, i, vp]),
        "dff_conv3d_scratch_bytes": (sz, [i, i, i, i, i]),
        "dff_conv3d": (i, [vp, i, vp, i, i, i, i, i, fp, i, i, i, i, i, i, i, fp, fp, vp, vp, i, vp, i, i, vp, i, vp]),
        "dff_conv3d_ex": (i, [vp, i, vp, i, i, i, i, i, fp, i, i, i, i, i, i, i, fp, fp, vp, vp, i, vp, i, i, i, vp, vp, fp, fp, i, i, vp,
                              i, vp]),
        "dff_to_pair_packed": (i, [fp, i, i, i, i, vp, i, vp]),
        "dff_depth_head": (i, [fp, i, i, fp, c.POINTER(i64), i, i, i, i, fp, i, vp]),
        "dff_srd_attention": (i, [vp, i, i, i, i, i, fp, fp, vp, vp, i, vp]),
        "dff_fov_warp": (i, [fp, fp, fp, i, i, i, i, i, fp, fp, i, vp]),
        "dff_fov_warp_cl": (i, [vp, fp, fp, i, i, i, i, i, vp, i, i, vp]),
        "dff_pair_volume": (i, [vp, fp, fp, i, i, i, i, i, vp, i, i, vp]),
        "dff_spatial_mean_accum": (i, [fp, i, i, i, i, i, fp, c.c_float, c.c_float, c.c_float, fp, i, vp]),
        "dff_to_channels_last": (i, [fp, i, i, i, i, i, vp, i, i, i, vp]),
        "dff_from_channels_last": (i, [vp, i, i, i, i, i, i, i, fp, i, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    return sig


def lib():
    """The loaded C-ABI library.  Raises if it has not been built (`python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise DffError("dff_b200: %s is missing — build it with __graft_entry__.build(); there is no "
                                   "CPU or eager fallback" % LIB_PATH)
                l = ctypes.CDLL(LIB_PATH)
                _declare(l)
                _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise DffError("dff_b200 error %d: %s" % (rc, lib().dff_last_error().decode()))


def default_precision():
    return os.environ.get("DFF_B200_PRECISION", "fp32")


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


_param_names = {}


def param_names(net=NET_DFF):
    names = _param_names.get(net)
    if names is None:
        l = lib()
        names = _param_names[net] = tuple(l.dff_param_name(net, i).decode() for i in range(l.dff_param_count(net)))
    return names


def _require_cuda(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise DffError("dff_b200: %s must be a CUDA tensor (the hot path has no CPU implementation)" % what)


_checked_devices = set()


def _check_device(index):
    if index not in _checked_devices:
        check(lib().dff_check_device(index))
        _checked_devices.add(index)


# ---------------------------------------------------------------------------------------------------------------
# weights
# ---------------------------------------------------------------------------------------------------------------
class PackedWeights:
    """Kernel-layout weight pack of one network, one cache slot per CUDA device, shared by reference with the
    `nn.DataParallel` replicas of its owner.

    `nn.DataParallel` (reference call sites Depth_Estimation_Test/test.py:30-32, train_codes/train_code_Defocus.py:63) rebuilds
    the replicas for every forward: their parameters are freshly broadcast plain attributes (`_parameters` is empty) and their
    `__dict__` is a shallow copy of the owner's.  So the cache lives on the OWNER (`owner_of`), the key is computed from the
    owner's tensors — `(data_ptr, _version)` of every parameter / running statistic plus the `_version` of every
    `num_batches_tracked` — and a slot is (re)packed from the calling replica's own tensors, which live on its device.
    Updates the key cannot see (`p.data.mul_()` style writes do not bump `_version`) need `invalidate()`; `module.train()` /
    `.eval()` transitions invalidate as well.
    """

    def __init__(self, net=NET_DFF):
        self.net = net
        self._lock = threading.Lock()
        self._slots = {}       # device index -> (key, raw, packed)
        self._paths = None     # [(submodule, dict name, leaf)] of the owner, in the library's parameter order
        self._nbt = None

    def __deepcopy__(self, memo):   # copies / pickles of a module start with an empty cache
        return PackedWeights(self.net)

    def __reduce__(self):
        return (PackedWeights, (self.net,))

    def invalidate(self):
        with self._lock:
            self._slots.clear()
            self._paths = None

    @staticmethod
    def _resolve(module, name):
        obj = module
        for part in name.split("."):
            try:
                obj = getattr(obj, part)
            except AttributeError:
                raise DffError("dff_b200: module has no tensor named %r (state_dict layout mismatch)" % name)
        if not isinstance(obj, torch.Tensor):
            raise DffError("dff_b200: %r is not a tensor (state_dict layout mismatch)" % name)
        return obj

    def _key(self, owner, device):
        if self._paths is None:
            paths, nbt = [], []
            for n in param_names(self.net):
                head, _, leaf = n.rpartition(".")
                try:
                    sub = owner.get_submodule(head) if head else owner
                except AttributeError:
                    raise DffError("dff_b200: module has no tensor named %r (state_dict layout mismatch)" % n)
                if leaf in sub._parameters:
                    paths.append((sub._parameters, leaf))
                elif leaf in sub._buffers:
                    paths.append((sub._buffers, leaf))
                else:
                    raise DffError("dff_b200: module has no tensor named %r (state_dict layout mismatch)" % n)
                if leaf == "running_var" and "num_batches_tracked" in sub._buffers:
                    nbt.append((sub._buffers, "num_batches_tracked"))
            self._paths, self._nbt = paths, nbt
        key = [device.index]
        for d, leaf in self._paths:
            t = d[leaf]
            key.append(t.data_ptr())
            key.append(t._version)
        for d, leaf in self._nbt:
            t = d[leaf]
            if t is not None:
                key.append(t._version)
        return tuple(key)

    def get(self, module, device, owner=None):
        """Packed weights for `module` (the owner itself or one of its DataParallel replicas) on `device`."""
        owner = owner if owner is not None else module
        with self._lock:
            try:
                key = self._key(owner, device)
            except (KeyError, AttributeError):   # a submodule / tensor was replaced since the paths were resolved
                self._paths = None
                key = self._key(owner, device)
            slot = self._slots.get(device.index)
            if slot is not None and slot[0] == key:
                return slot[2]
        # pack outside the lock (replica threads of different devices pack concurrently); last writer wins, both are valid
        l = lib()
        with torch.no_grad():
            ts = [self._resolve(module, n) for n in param_names(self.net)]
            raw = torch.cat([t.detach().reshape(-1).to(device=device, dtype=torch.float32) for t in ts])
        if raw.numel() != l.dff_raw_numel(self.net):
            raise DffError("dff_b200: raw parameter count %d != %d" % (raw.numel(), l.dff_raw_numel(self.net)))
        packed = torch.empty(l.dff_packed_bytes(self.net), dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            check(l.dff_pack_weights(self.net, _ptr(raw), _ptr(packed), device.index, _stream(device)))
        with self._lock:
            self._slots[device.index] = (key, raw, packed)
        return packed


_cache_lock = threading.Lock()


def owner_of(module):
    """The module whose parameters `module` mirrors: itself, or the source of an `nn.DataParallel` replica
    (`PackedOwnerMixin._replicate_for_data_parallel` records it)."""
    return module.__dict__.get("_dff_source") or module


def packed_cache(module, net=NET_DFF):
    owner = owner_of(module)
    cache = owner.__dict__.get("_dff_packed")
    if cache is None:
        with _cache_lock:
            cache = owner.__dict__.get("_dff_packed")
            if cache is None:
                cache = PackedWeights(net)
                owner.__dict__["_dff_packed"] = cache
    return cache


def packed_weights(module, device, net=NET_DFF):
    return packed_cache(module, net).get(module, device, owner_of(module))


class PackedOwnerMixin:
    """For the drop-in modules that own a weight pack: replicas remember their source, mode flips and explicit calls invalidate."""

    def _replicate_for_data_parallel(self):
        replica = super()._replicate_for_data_parallel()
        replica.__dict__["_dff_source"] = owner_of(self)
        return replica

    def train(self, mode=True):
        if mode != self.training:
            self.invalidate_packed_weights()
        return super().train(mode)

    def invalidate_packed_weights(self):
        """Call after writes the cache key cannot see (`p.data.copy_()`, raw-pointer updates)."""
        cache = owner_of(self).__dict__.get("_dff_packed")
        if cache is not None:
            cache.invalidate()


_ws_lock = threading.Lock()
_ws_cache = {}


def _workspace(device, B, S, H, W, mode):
    """One workspace per device (the largest shape seen so far is kept; a forward on a device is stream-ordered)."""
    n = lib().dff_workspace_bytes(B, S, H, W, mode)
    if n == 0:
        check(-1)
    with _ws_lock:
        ws = _ws_cache.get(device.index)
        if ws is None or ws.numel() < n:
            _ws_cache.pop(device.index, None)
            ws = torch.empty(n, dtype=torch.uint8, device=device)
            _ws_cache[device.index] = ws
    return ws


def _mode(net):
    prec = getattr(net, "precision", "fp32")
    if prec not in ("fp32", "bf16"):
        raise DffError("dff_b200: precision must be 'fp32' or 'bf16'")
    if prec == "bf16" and os.environ.get("DFF_B200_NO_TC") == "1":
        return BF16 | NO_TC   # debugging aid: bf16 storage, FFMA kernels
    if prec == "bf16" and os.environ.get("DFF_B200_NO_SLAB") == "1":
        return BF16 | NO_SLAB
    return BF16 if prec == "bf16" else FP32


# ---------------------------------------------------------------------------------------------------------------
# DFF_net forward (reference train_codes/Depth_Estimation_Network.py:77-137)
# ---------------------------------------------------------------------------------------------------------------
def _any_bn_training(net):
    bns = net.__dict__.get("_dff_bns")
    if bns is None:
        bns = net.__dict__["_dff_bns"] = [m for m in net.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
    return any(m.training for m in bns)


def _any_requires_grad(net):
    # (DataParallel replicas hold their parameters as plain attributes: ask the module they mirror)
    return any(p.requires_grad for p in owner_of(net).parameters())


def _pad32(n):
    return (n + 31) // 32 * 32


def _fd_view(focus_dists, dev, B, S, H, W):
    fd = focus_dists.to(dev) if dev is not None else focus_dists
    while fd.dim() < 4:
        fd = fd.unsqueeze(0)
    return fd.expand(B, S, H, W)


def dff_net_forward(net, FS, focus_dists, return_costs=False):
    """`DFF_net.forward(FS, focus_dists)`.  FS is the reference's fp32 (B,3,S,H,W) tensor — or, as an extension for input
    staging, the uint8 (B,S,H0,W0,3) stacks the datasets store: normalisation, -1 padding to multiples of 32
    and the layout change then happen on the GPU, and the (B,H,W) maps come back padded (callers crop `[:H0,:W0]` exactly as
    Depth_Estimation_Test/test.py:125 does)."""
    _require_cuda(FS, "FS")
    _require_cuda(focus_dists, "focus_dists")
    u8 = FS.dtype == torch.uint8
    if u8:
        if FS.dim() != 5 or FS.shape[4] != 3:
            raise DffError("dff_b200: uint8 FS must be (B,S,H,W,3) as the datasets store it, got %s" % (tuple(FS.shape),))
    elif FS.dim() != 5 or FS.shape[1] != 3:
        raise DffError("dff_b200: FS must be (B,3,S,H,W), got %s" % (tuple(FS.shape),))
    if (not u8 and FS.dtype != torch.float32) or focus_dists.dtype != torch.float32:
        raise DffError("dff_b200: FS and focus_dists must be float32 (as the reference dataloaders produce)")
    # BatchNorm behaviour follows the module's mode (batch statistics iff training, per BatchNorm3d as torch does); the autograd
    # tape is needed iff gradients are enabled and something on the path requires them.  Only "eval statistics, no tape" is the
    # single-call inference path; everything else runs operator by operator (train.py).
    bn_batch = net.training or _any_bn_training(net)
    need_tape = torch.is_grad_enabled() and (FS.requires_grad or _any_requires_grad(net))
    if bn_batch or need_tape:
        if return_costs:
            raise DffError("dff_b200: return_costs is an eval-mode debugging aid")
        from . import train as _train
        if u8:
            FS = stage_u8(FS)
        return _train.dff_net_train_forward(net, FS, focus_dists)
    dev = FS.device
    _check_device(dev.index)
    if u8:
        B, S, H0, W0, _ = FS.shape
        H, W = _pad32(H0), _pad32(W0)
    else:
        B, _, S, H, W = FS.shape
    FS = FS.contiguous()
    fd = _fd_view(focus_dists, dev, B, S, H, W)
    strides = (ctypes.c_int64 * 4)(*fd.stride())
    mode = _mode(net)
    packed = packed_weights(net, dev)
    ws = _workspace(dev, B, S, H, W, mode)
    outs = [torch.empty((B, H, W), dtype=torch.float32, device=dev) for _ in range(4)]
    out_ptrs = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in outs])
    costs, cost_ptrs = None, None
    if return_costs:
        costs = [torch.empty((B, S, H // r, W // r), dtype=torch.float32, device=dev) for r in (8, 4, 2, 1)]
        cost_ptrs = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in costs])
    with torch.cuda.device(dev):
        if u8:
            check(lib().dff_forward_u8(_ptr(packed), _ptr(FS), H0, W0, _ptr(fd), strides, B, S, H, W, out_ptrs, cost_ptrs, _ptr(ws),
                                       ws.numel(), mode, dev.index, _stream(dev)))
        else:
            check(lib().dff_forward(_ptr(packed), _ptr(FS), _ptr(fd), strides, B, S, H, W, out_ptrs, cost_ptrs, _ptr(ws),
                                    ws.numel(), mode, dev.index, _stream(dev)))
    if return_costs:
        return tuple(outs), tuple(costs)
    return tuple(outs)


def stage_u8(FS_u8):
    """The dataloader tail on the GPU (reference test_Dataloader.py:122-141): uint8 (B,S,H0,W0,3) -> fp32 (B,3,S,H,W),
    `x/127.5 - 1`, H and W padded to multiples of 32 with -1."""
    _require_cuda(FS_u8, "FS")
    B, S, H0, W0, _ = FS_u8.shape
    H, W = _pad32(H0), _pad32(W0)
    out = torch.empty((B, 3, S, H, W), dtype=torch.float32, device=FS_u8.device)
    with torch.cuda.device(FS_u8.device):
        check(lib().dff_stage_u8(_ptr(FS_u8.contiguous()), H0, W0, B, S, H, W, _ptr(out), FS_u8.device.index, _stream(FS_u8.device)))
    return out


def forward_host(net, FS_host, fd_host, device, micro_batch=16, outputs=(True, True, True, True)):
    """The eval loop's upload + forward + download as ONE library call on HOST tensors (pinned for full speed):
    `dff_forward_host` for fp32 (B,3,S,H,W) stacks, `dff_forward_host_u8` for uint8 (B,S,H0,W0,3) stacks.  Returns the four
    (B,H,W) maps as pinned CPU tensors (None where `outputs[j]` is False — test.py:118-121 reads only pred3)."""
    if net.training or FS_host.is_cuda or fd_host.is_cuda:
        raise DffError("dff_b200: forward_host is the eval path on host tensors")
    dev = torch.device(device)
    _check_device(dev.index)
    u8 = FS_host.dtype == torch.uint8
    if u8:
        B, S, H0, W0, _ = FS_host.shape
        H, W = _pad32(H0), _pad32(W0)
    else:
        B, _, S, H, W = FS_host.shape
        H0, W0 = H, W
    l = lib()
    FS_host = FS_host.contiguous()
    fd = _fd_view(fd_host, None, B, S, H, W)
    if not u8:
        fd = fd.contiguous()
    elif any(st < 0 for st in fd.stride()):
        fd = fd.contiguous()
    strides = (ctypes.c_int64 * 4)(*fd.stride())
    mode = _mode(net)
    mb = max(1, min(micro_batch, B))
    packed = packed_weights(net, dev)
    ws = _workspace(dev, mb, S, H, W, mode)
    nio = l.dff_host_io_bytes_u8(mb, S, H0, W0, H, W, strides) if u8 else l.dff_host_io_bytes(mb, S, H, W)
    with _ws_lock:
        io = _io_cache.get(dev.index)
        if io is None or io.numel() < nio:
            _io_cache.pop(dev.index, None)
            io = torch.empty(nio, dtype=torch.uint8, device=dev)
            _io_cache[dev.index] = io
    outs = [torch.empty((B, H, W), dtype=torch.float32).pin_memory() if want else None for want in outputs]
    hp = (ctypes.c_void_p * 4)(*[o.data_ptr() if o is not None else None for o in outs])
    with torch.cuda.device(dev):
        if u8:
            check(l.dff_forward_host_u8(_ptr(packed), _ptr(FS_host), H0, W0, _ptr(fd), strides, B, mb, S, H, W, hp, _ptr(io), _ptr(ws),
                                        ws.numel(), mode, dev.index, _stream(dev)))
        else:
            check(l.dff_forward_host(_ptr(packed), _ptr(FS_host), _ptr(fd), strides, B, mb, S, H, W, hp, _ptr(io), _ptr(ws),
                                     ws.numel(), mode, dev.index, _stream(dev)))
    return tuple(outs)


_io_cache = {}


# ---------------------------------------------------------------------------------------------------------------
# single operators (unit-parity surface of the C-ABI; tensors in the reference's (B,C,S,H,W) layout)
# ---------------------------------------------------------------------------------------------------------------
def _elem(bf16):
    return BF16 if bf16 else FP32


def to_channels_last(x, Cp=None, bf16=False):
    """(B,C,S,H,W) fp32 -> channels-last (B,S,H,W,Cp) fp32|bf16 via the library's layout kernel."""
    _require_cuda(x, "x")
    B, C, S, H, W = x.shape
    Cp = Cp or (C + 3) // 4 * 4
    out = torch.empty((B, S, H, W, Cp), dtype=torch.bfloat16 if bf16 else torch.float32, device=x.device)
    check(lib().dff_to_channels_last(_ptr(x.contiguous()), B, C, S, H, W, _ptr(out), Cp, _elem(bf16), x.device.index,
                                     _stream(x.device)))
    return out


def from_channels_last(x, C=None):
    _require_cuda(x, "x")
    B, S, H, W, Cp = x.shape
    C = C or Cp
    out = torch.empty((B, C, S, H, W), dtype=torch.float32, device=x.device)
    check(lib().dff_from_channels_last(_ptr(x.contiguous()), B, C, S, H, W, Cp, _elem(x.dtype == torch.bfloat16), _ptr(out),
                                       x.device.index, _stream(x.device)))
    return out


def conv3d(x, weight, stride_hw=1, dil_hw=1, transposed=False, scale=None, shift=None, res_pre=None, res_post=None,
           relu=False, x2=None, bf16=False, tensor_cores=False):
    """One fused conv call on reference-layout tensors.  x (B,C0,S,H,W) [x2 (B,C1,S,H,W) = virtual concat];
    weight in the reference layout; res_* (B,Cout,S,OH,OW).  Returns (B,Cout,S,OH,OW) fp32."""
    l = lib()
    dev = x.device
    B, C0, S, IH, IW = x.shape
    Cout = weight.shape[1] if transposed else weight.shape[0]
    kd, kh, kw = weight.shape[2:]
    cal = 8 if tensor_cores else 4
    C0p = (C0 + cal - 1) // cal * cal
    a = to_channels_last(x, C0p, bf16)
    b, C1p = None, 0
    w = weight.detach().to(dev, torch.float32)
    if x2 is not None:
        C1 = x2.shape[1]
        C1p = (C1 + cal - 1) // cal * cal
        b = to_channels_last(x2, C1p, bf16)
    if C0p != C0 or (x2 is not None and C1p != x2.shape[1]):
        # zero weights for the padded input channels
        cin_axis = 0 if transposed else 1
        parts = [w.narrow(cin_axis, 0, C0), w.new_zeros(_with(w.shape, cin_axis, C0p - C0))]
        if x2 is not None:
            parts += [w.narrow(cin_axis, C0, x2.shape[1]), w.new_zeros(_with(w.shape, cin_axis, C1p - x2.shape[1]))]
        w = torch.cat(parts, cin_axis)
    w = w.contiguous()
    OH, OW = (IH * 2, IW * 2) if transposed else (IH // stride_hw, IW // stride_hw)
    Cop = Cout if Cout % 8 == 0 else Cout  # stored output channels = Cout (scalar store path when not a multiple of 8)
    out = torch.empty((B, S, OH, OW, Cop), dtype=torch.bfloat16 if bf16 else torch.float32, device=dev)
    rp = to_channels_last(res_pre, Cop, bf16) if res_pre is not None else None
    rq = to_channels_last(res_post, Cop, bf16) if res_post is not None else None
    sc = scale.detach().to(dev, torch.float32).contiguous() if scale is not None else None
    sh = shift.detach().to(dev, torch.float32).contiguous() if shift is not None else None
    if sc is not None and sc.numel() % 16:
        sc = torch.cat([sc, sc.new_ones(16 - sc.numel() % 16)])
    if sh is not None and sh.numel() % 16:
        sh = torch.cat([sh, sh.new_zeros(16 - sh.numel() % 16)])
    scratch = torch.empty(l.dff_conv3d_scratch_bytes(C0p + C1p, Cout, kd, kh, kw), dtype=torch.uint8, device=dev)
    check(l.dff_conv3d(_ptr(a), C0p, _ptr(b), C1p, B, S, IH, IW, _ptr(w), Cout, kd, kh, kw, stride_hw, dil_hw,
                       1 if transposed else 0, _ptr(sc), _ptr(sh), _ptr(rp), _ptr(rq), 1 if relu else 0, _ptr(out),
                       _elem(bf16), int(tensor_cores), _ptr(scratch), dev.index, _stream(dev)))
    return from_channels_last(out, Cout)


def conv3d_forward_plan(x, weight, stride_hw=1, dil_hw=1, transposed=False, scale=None, shift=None, res_pre=None, res_post=None,
                        relu=False, x2=None, aux_add=None, proj_w=None, proj_on_aux=False, skip_out=False):
    """One conv exactly as `dff_forward` would run a layer of this shape in bf16 (plan 4 of `dff_conv3d_ex`: folded forms, second
    output, fused classifier).  Reference-layout fp32 tensors in, dict of reference-layout fp32 tensors out:
    out (B,Cout,S,OH,OW) [unless skip_out], aux = out + aux_add, proj (B,S,OH,OW)."""
    l = lib()
    dev = x.device
    B, Cx, S, IH, IW = x.shape
    Cout = weight.shape[1] if transposed else weight.shape[0]
    kd, kh, kw = weight.shape[2:]
    w = weight.detach().to(dev, torch.float32).contiguous()
    pair = (not transposed) and Cx == 3 and (kd, kh, kw) == (1, 9, 9) and dil_hw == 2
    if pair:
        a = torch.empty((B, S, IH, IW + 2, 8), dtype=torch.bfloat16, device=dev)
        check(l.dff_to_pair_packed(_ptr(x.contiguous()), B, S, IH, IW, _ptr(a), dev.index, _stream(dev)))
        C0p, IWs = 8, IW + 2
    else:
        if Cx % 8:
            raise DffError("conv3d_forward_plan: channel counts must be multiples of 8")
        a, C0p, IWs = to_channels_last(x, Cx, True), Cx, IW
    b, C1p = (to_channels_last(x2, x2.shape[1], True), x2.shape[1]) if x2 is not None else (None, 0)
    OH, OW = (IH * 2, IW * 2) if transposed else (IH // stride_hw, IW // stride_hw)
    out = torch.zeros((B, S, OH, OW, Cout), dtype=torch.bfloat16, device=dev)
    cl = lambda t: to_channels_last(t, Cout, True) if t is not None else None
    rp, rq, ax = cl(res_pre), cl(res_post), cl(aux_add)
    aux = torch.empty_like(out) if ax is not None else None
    pw = proj_w.detach().to(dev, torch.float32).contiguous() if proj_w is not None else None
    pout = torch.empty((B, S, OH, OW), dtype=torch.float32, device=dev) if pw is not None else None
    sc = scale.detach().to(dev, torch.float32).contiguous() if scale is not None else None
    sh = shift.detach().to(dev, torch.float32).contiguous() if shift is not None else None
    scratch = torch.empty(l.dff_conv3d_scratch_bytes(max(C0p + C1p, 8), Cout, kd, kh, kw), dtype=torch.uint8, device=dev)
    check(l.dff_conv3d_ex(_ptr(a), C0p, _ptr(b), C1p, B, S, IH, IWs, _ptr(w), Cout, kd, kh, kw, stride_hw, dil_hw, 1 if transposed else 0,
                          _ptr(sc), _ptr(sh), _ptr(rp), _ptr(rq), 1 if relu else 0, _ptr(out), BF16, 4, 1 if pair else 0, _ptr(ax),
                          _ptr(aux), _ptr(pw), _ptr(pout), 1 if proj_on_aux else 0, 1 if skip_out else 0, _ptr(scratch), dev.index,
                          _stream(dev)))
    res = {}
    if not skip_out:
        res["out"] = from_channels_last(out, Cout)
    if aux is not None:
        res["aux"] = from_channels_last(aux, Cout)
    if pout is not None:
        res["proj"] = pout
    return res


def _with(shape, axis, n):
    s = list(shape)
    s[axis] = n
    return s


def srd_attention(x, w_a, w_b):
    """SRD channel-attention branch (reference Depth_Estimation_Network.py:399-407) in one pass on bf16 channels-last data:
    x (B,C,S,H,W) fp32, w_a (C,C,3,1,1), w_b (C,C,1,1,1) -> x + relu(conv1x1x1(relu(conv3x1x1(x)))) as (B,C,S,H,W) fp32."""
    _require_cuda(x, "x")
    B, C, S, H, W = x.shape
    a = to_channels_last(x, C, True)
    out = torch.empty_like(a)
    scratch = torch.empty(4 * C * C, dtype=torch.float32, device=x.device)
    check(lib().dff_srd_attention(_ptr(a), B, S, H, W, C, _ptr(w_a.detach().to(x.device, torch.float32).contiguous()),
                                  _ptr(w_b.detach().to(x.device, torch.float32).contiguous()), _ptr(out), _ptr(scratch),
                                  x.device.index, _stream(x.device)))
    return from_channels_last(out, C)


def depth_head(cost, focus_dists, H, W):
    """cost (B,S,h,w) fp32, focus_dists broadcastable to (B,S,H,W) -> (B,H,W)."""
    _require_cuda(cost, "cost")
    B, S, h, w = cost.shape
    fd = focus_dists.to(cost.device)
    while fd.dim() < 4:
        fd = fd.unsqueeze(0)
    fd = fd.expand(B, S, H, W)
    strides = (ctypes.c_int64 * 4)(*fd.stride())
    out = torch.empty((B, H, W), dtype=torch.float32, device=cost.device)
    check(lib().dff_depth_head(_ptr(cost.contiguous()), h, w, _ptr(fd), strides, B, S, H, W, _ptr(out), cost.device.index,
                               _stream(cost.device)))
    return out


def fov_warp(x, alpha, fovs, want_flow=True):
    """FlowNetwork.FOV_warp on (B,C,S,H,W) fp32; alpha (B,3,S,1,1) or None; fovs (B,1,S,1,1)."""
    _require_cuda(x, "x")
    B, C, S, H, W = x.shape
    out = torch.empty_like(x)
    flow = torch.empty((B, 2, S, H, W), dtype=torch.float32, device=x.device) if want_flow else None
    al = alpha.reshape(B, 3, S).contiguous().float() if alpha is not None else None
    fv = fovs.reshape(B, S).contiguous().float()
    check(lib().dff_fov_warp(_ptr(x.contiguous()), _ptr(al), _ptr(fv), B, C, S, H, W, _ptr(out), _ptr(flow), x.device.index,
                             _stream(x.device)))
    return out, flow
