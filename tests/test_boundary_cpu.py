"""Host-side logic that needs no GPU: the drop-in module's state_dict contract, seed-identical construction,
the C-ABI library's exports and its parameter handshake, and the refusal to run on the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, golden


def _net():
    from dffinthewild_b200.Depth_Estimation_Network import Network
    torch.manual_seed(0)
    return Network()


def test_state_dict_layout_matches_reference():
    lay = golden("state_layout.npz")
    sd = _net().state_dict()
    assert list(sd.keys()) == [str(k) for k in lay["keys"]]
    assert len(sd) == 384
    for (k, v), shp, dt in zip(sd.items(), lay["shapes"], lay["dtypes"]):
        assert str(tuple(v.shape)) == str(shp), k
        assert str(v.dtype) == str(dt), k


def test_seeded_construction_is_reference_identical():
    lay = golden("state_layout.npz")
    sd = _net().state_dict()
    for v, s, a in zip(sd.values(), lay["seed0_sum"], lay["seed0_abs"]):
        assert abs(float(v.double().sum()) - s) <= 1e-9 * max(1.0, abs(a))
        assert abs(float(v.double().abs().sum()) - a) <= 1e-9 * max(1.0, abs(a))


def test_call_site_surface():
    """The model-handling lines of test.py:30-32,77-85 and train_code_DDFF.py:48,62-67 work unchanged."""
    net = _net().cpu()
    dp = torch.nn.DataParallel(net)
    sd = dp.module.state_dict()
    dp.module.load_state_dict(sd, strict=True)
    dp.load_state_dict(dp.state_dict(), strict=True)
    assert all(k.startswith("module.DFF_net.") for k in dp.state_dict())
    opt = torch.optim.Adam(dp.parameters(), lr=1e-4, betas=(0.9, 0.99))
    assert sum(p.numel() for g in opt.param_groups for p in g["params"]) == 4038832
    dp.eval(); dp.train()
    assert hasattr(net, "DFF_net")


def test_cpu_forward_is_refused(built_lib):
    from dffinthewild_b200.runtime import DffError
    net = _net().eval()
    with pytest.raises(DffError):
        net(torch.zeros(1, 3, 2, 32, 32), torch.zeros(1, 2, 32, 32))


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "dff_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(dff_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 18
    lib = ctypes.CDLL(built_lib)
    for n in names:
        assert hasattr(lib, n), n
    lib.dff_abi_version.restype = ctypes.c_int
    assert lib.dff_abi_version() == 1


def test_parameter_handshake(built_lib):
    from dffinthewild_b200 import runtime as rt
    sd = _net().DFF_net.state_dict()
    names = rt.param_names(rt.NET_DFF)
    l = rt.lib()
    assert len(names) == len(set(names))
    off = 0
    for i, n in enumerate(names):
        assert n in sd, n
        assert sd[n].numel() == l.dff_param_numel(rt.NET_DFF, i), n
        assert l.dff_param_offset(rt.NET_DFF, i) == off
        off += sd[n].numel()
    assert off == l.dff_raw_numel(rt.NET_DFF)
    # everything the forward never touches: the 12 dead parameters, their BN buffers, and num_batches_tracked
    unused = [k for k in sd if k not in names and not k.endswith("num_batches_tracked")]
    assert all(("redir3" in k) or ("pre_conv" in k) for k in unused), unused
    assert len([k for k in unused if "running" not in k]) == 12


def test_workspace_size_and_shape_errors(built_lib):
    from dffinthewild_b200 import runtime as rt
    l = rt.lib()
    a = l.dff_workspace_bytes(1, 10, 224, 224, rt.FP32)
    b = l.dff_workspace_bytes(2, 10, 224, 224, rt.FP32)
    h = l.dff_workspace_bytes(1, 10, 224, 224, rt.BF16)
    assert a > 0 and abs(b - 2 * a) <= 1 << 16 and h < a
    assert l.dff_workspace_bytes(1, 10, 100, 224, rt.FP32) == 0  # H not a multiple of 32
    assert b"multiples of 32" in l.dff_last_error()


def test_operator_plan_dry_run(built_lib):
    """The library plans a forward without a GPU (`dff_forward_profiled` with null pointers): the operator list is what bench.py's
    roofline divides by — its algorithmic FLOPs must be SURVEY.md §8(d)'s 93,563 FLOP/voxel (2*MACs of the 70 executed conv layers,
    classifier and attention products included) and the bf16 plan must stay at one launch per fused operator."""
    import ctypes
    from dffinthewild_b200 import runtime as rt
    l = rt.lib()
    N = 256
    ms, fl, by = (ctypes.c_float * N)(), (ctypes.c_double * N)(), (ctypes.c_double * N)()
    la, nm, n = (ctypes.c_int * N)(), ctypes.create_string_buffer(N * 64), ctypes.c_int(0)
    B, S, H, W = 2, 10, 384, 576
    for mode, max_launches in ((rt.BF16, 80), (rt.FP32, 140)):
        rc = l.dff_forward_profiled(None, None, None, None, B, S, H, W, None, None, 0, mode, 0, None, N, ms, fl, by, la, nm, ctypes.byref(n))
        assert rc == 0, l.dff_last_error()
        names = [nm.raw[k * 64:(k + 1) * 64].split(b"\0")[0].decode() for k in range(n.value)]
        assert names[0] == "to_channels_last" and names[-1] == "depth_heads"
        assert "dres4.conv6.0" in names and "FM_measure.Focus_extraction.0.0" in names
        flops = sum(fl[k] for k in range(n.value))
        vox = B * S * H * W
        assert abs(flops / vox - 93563.0) / 93563.0 < 0.01, flops / vox
        launches = sum(la[k] for k in range(n.value))
        assert n.value <= launches <= max_launches, launches
        assert all(by[k] > 0 for k in range(n.value))


def _fake_replica(net):
    """What torch.nn.parallel.replicate builds for one device (torch/nn/parallel/replicate.py), without needing a GPU:
    per-module `_replicate_for_data_parallel()` shells, children re-linked, parameters re-attached as plain tensors."""
    mods = list(net.modules())
    idx = {m: i for i, m in enumerate(mods)}
    reps = [m._replicate_for_data_parallel() for m in mods]
    for m, r in zip(mods, reps):
        for k, child in m._modules.items():
            r._modules[k] = None if child is None else reps[idx[child]]
        for k, p in m._parameters.items():
            if p is not None:
                t = p.detach().clone()      # broadcast copy: a plain tensor, not an nn.Parameter
                setattr(r, k, t)
        for k, b in m._buffers.items():
            r._buffers[k] = None if b is None else b.clone()
    return reps[0]


def test_dataparallel_replica_resolves_every_tensor(built_lib):
    """nn.DataParallel replicas have empty `_parameters` (weights are plain attributes) and a shallow copy of the owner's
    `__dict__`: the weight pack must find every tensor by attribute path and keep its cache on the owner
    (reference call sites Depth_Estimation_Test/test.py:30-32, train_codes/train_code_Defocus.py:63)."""
    from dffinthewild_b200 import runtime as rt
    net = _net()
    rep = _fake_replica(net)
    assert len(list(rep.DFF_net.named_parameters())) == 0           # the situation round 1 broke on
    assert rt.owner_of(rep.DFF_net) is net.DFF_net
    assert rt.owner_of(net.DFF_net) is net.DFF_net
    for n in rt.param_names(rt.NET_DFF):
        a, b = rt.PackedWeights._resolve(rep.DFF_net, n), rt.PackedWeights._resolve(net.DFF_net, n)
        assert a.data_ptr() != b.data_ptr() and torch.equal(a, b), n
    assert rt.packed_cache(rep.DFF_net) is rt.packed_cache(net.DFF_net)
    assert rt._any_requires_grad(rep.DFF_net)


def test_weight_cache_key_and_invalidation(built_lib):
    import copy
    from dffinthewild_b200 import runtime as rt
    net = _net().DFF_net
    cache = rt.packed_cache(net)
    dev = torch.device("cuda", 0)   # (only the index is used by the key)
    k0 = cache._key(net, dev)
    assert cache._key(net, dev) == k0
    with torch.no_grad():
        net.classif3[0].weight.mul_(2.0)
    k1 = cache._key(net, dev)
    assert k1 != k0
    bn = net.dres4.conv2[1]
    bn.num_batches_tracked += 1            # what a train-mode forward does
    k2 = cache._key(net, dev)
    assert k2 != k1
    torch.autograd.graph.increment_version(bn.running_mean)   # what train.py does after the library updated the buffer in place
    assert cache._key(net, dev) != k2
    # mode flips invalidate; deep copies and pickles start empty and do not share the lock
    cache._slots[0] = ("k", None, None)
    net.eval()
    assert not cache._slots
    cache._slots[0] = ("k", None, None)
    net.invalidate_packed_weights()
    assert not cache._slots
    twin = copy.deepcopy(net)
    assert rt.packed_cache(twin) is not cache and not rt.packed_cache(twin)._slots
    import pickle
    assert isinstance(pickle.loads(pickle.dumps(cache)), rt.PackedWeights)


def test_dispatch_follows_mode_and_grad(built_lib, monkeypatch):
    """BatchNorm behaviour follows module.training; the tape follows grad mode.  Only eval + no tape is the single-call path."""
    from dffinthewild_b200 import runtime as rt
    from dffinthewild_b200 import train as tr
    seen = []
    monkeypatch.setattr(rt, "_require_cuda", lambda t, what: None)
    monkeypatch.setattr(tr, "dff_net_train_forward", lambda net, FS, fd: seen.append("train") or ())
    monkeypatch.setattr(rt, "_check_device", lambda i: (_ for _ in ()).throw(rt.DffError("inference path")))
    net = _net().DFF_net
    FS, fd = torch.zeros(1, 3, 2, 32, 32), torch.zeros(1, 2, 32, 32)

    def path(**kw):
        seen.clear()
        try:
            rt.dff_net_forward(net, FS, fd)
        except rt.DffError as e:
            assert "inference path" in str(e)
            return "infer"
        return seen[0]

    net.train()
    assert path() == "train"
    with torch.no_grad():
        assert path() == "train"              # batch statistics + running-stat updates, no tape
    net.eval()
    assert path() == "train"                  # eval statistics, but parameters require grad: differentiable outputs
    with torch.no_grad():
        assert path() == "infer"
    for p in net.parameters():
        p.requires_grad_(False)
    assert path() == "infer"
    FS.requires_grad_(True)
    assert path() == "train"                  # input gradients of a frozen eval network
    FS.requires_grad_(False)
    net.dres2.conv2[1].train()
    assert path() == "train"                  # one BatchNorm3d in batch-statistics mode


def test_grad_bucket_survives_zero_grad():
    from dffinthewild_b200 import distributed as D
    net = _net()
    skip = D.unused_parameter_names(net)
    bucket = D.GradBucket(net, skip=skip)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    opt.zero_grad()                            # set_to_none=True: drops the views
    p0, p1 = bucket.params[0], bucket.params[1]
    assert p0.grad is None
    p0.grad = torch.full_like(p0, 3.0)         # autograd allocates a fresh gradient
    bucket.rebind()
    assert p0.grad.data_ptr() == bucket.flat.data_ptr() and float(bucket.flat[0]) == 3.0
    assert p1.grad is not None and float(p1.grad.abs().sum()) == 0.0
