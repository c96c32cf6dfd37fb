"""Host-side mirror of the reference's GCE + TRL modules on top of the C ABI.

Reference interfaces kept (same attribute names => same state_dict keys, same call signatures):
  Backbone(height, width).forward(x, b, t) -> (x_uncorr, x_corr, corr_map)   reid/models/basebranch.py:21-68
  BasicBlock(inplanes, planes)                                                reid/models/grl_model.py:51-85
  TRLBlock(feat_num).forward(x_uncorr, x_corr) -> (f_uncorr, f_corr)          reid/models/grl_model.py:87-180
  ResNet50_GRL_Model(...).forward(inputs, training=True) -> (x_uncorr, x_corr) reid/models/grl_model.py:184-228
  resnet50_grl(*args, **kwargs)                                               reid/models/grl_model.py:231-232

All head math runs in libgrl_b200.so (sm_100a).  There is no PyTorch/CPU fallback: on a machine
without the CUDA library or a B200 the forward raises.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn
from torch.nn import functional as F

from . import _lib

TP = "temporal_learning_block."
_DIRS = (("forward", "foreward"), ("backward", "backward"))


def head_param_names():
    """state_dict keys (relative to ResNet50_GRL_Model) consumed by the head, in a fixed order."""
    names = ["backbone.glo_fc.0.weight", "backbone.glo_fc.0.bias", "backbone.glo_fc.1.weight", "backbone.glo_fc.1.bias",
             "backbone.corr_atte.0.weight", "backbone.corr_atte.1.weight", "backbone.corr_atte.1.bias",
             "backbone.corr_atte.2.weight", "backbone.corr_atte.3.weight", "backbone.corr_atte.3.bias",
             "backbone.corr_atte.5.weight", "backbone.corr_atte.6.weight", "backbone.corr_atte.6.bias"]
    for direction, atte in _DIRS:
        m = TP + "uncorr_memo_" + direction
        names += [TP + direction + "_f1.0.weight", TP + direction + "_f1.0.bias",
                  TP + direction + "_f2.0.weight", TP + direction + "_f2.0.bias",
                  TP + "channel_atte_" + atte + "_corr.0.weight", TP + "channel_atte_" + atte + "_corr.2.weight",
                  m + ".conv1.weight", m + ".bn1.weight", m + ".bn1.bias",
                  m + ".conv2.weight", m + ".bn2.weight", m + ".bn2.bias",
                  m + ".conv3.weight", m + ".bn3.weight", m + ".bn3.bias"]
    return names


def gce_param_names():
    return [k for k in head_param_names() if k.startswith("backbone.")]


def trl_param_names():
    return [k for k in head_param_names() if k.startswith(TP)]


def head_buffer_names():
    bn = ["backbone.glo_fc.1", "backbone.corr_atte.1", "backbone.corr_atte.3", "backbone.corr_atte.6"]
    for direction, _ in _DIRS:
        bn += [TP + "uncorr_memo_" + direction + ".bn%d" % k for k in (1, 2, 3)]
    return bn


def _bn(sd, prefix):
    r = _lib.BnParams()
    r.weight = sd[prefix + ".weight"].data_ptr()
    r.bias = sd[prefix + ".bias"].data_ptr()
    r.running_mean = sd[prefix + ".running_mean"].data_ptr()
    r.running_var = sd[prefix + ".running_var"].data_ptr()
    return r


def pack_params(sd, which=3) -> _lib.HeadParams:
    """Fill the C parameter block with device pointers of fp32, contiguous CUDA tensors.
    which: 1 = GCE members only, 2 = TRL members only, 3 = both (the other members stay NULL)."""
    names = (gce_param_names() if which & 1 else []) + (trl_param_names() if which & 2 else [])
    for k in names:
        t = sd[k]
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError("grl_b200 head: parameter %s must be a contiguous float32 CUDA tensor" % k)
    p = _lib.HeadParams()
    if which & 1:
        _pack_gce_params(p, sd)
    if which & 2:
        _pack_trl_params(p, sd)
    return p


def _pack_gce_params(p, sd):
    p.glo_fc_w = sd["backbone.glo_fc.0.weight"].data_ptr()
    p.glo_fc_b = sd["backbone.glo_fc.0.bias"].data_ptr()
    p.glo_bn = _bn(sd, "backbone.glo_fc.1")
    p.atte0_w = sd["backbone.corr_atte.0.weight"].data_ptr()
    p.atte_bn1 = _bn(sd, "backbone.corr_atte.1")
    p.atte2_w = sd["backbone.corr_atte.2.weight"].data_ptr()
    p.atte_bn3 = _bn(sd, "backbone.corr_atte.3")
    p.atte5_w = sd["backbone.corr_atte.5.weight"].data_ptr()
    p.atte_bn6 = _bn(sd, "backbone.corr_atte.6")


def _pack_trl_params(p, sd):
    for d, (direction, atte) in enumerate(_DIRS):
        m = TP + "uncorr_memo_" + direction
        p.f1_w[d] = sd[TP + direction + "_f1.0.weight"].data_ptr()
        p.f1_b[d] = sd[TP + direction + "_f1.0.bias"].data_ptr()
        p.f2_w[d] = sd[TP + direction + "_f2.0.weight"].data_ptr()
        p.f2_b[d] = sd[TP + direction + "_f2.0.bias"].data_ptr()
        p.se1_w[d] = sd[TP + "channel_atte_" + atte + "_corr.0.weight"].data_ptr()
        p.se2_w[d] = sd[TP + "channel_atte_" + atte + "_corr.2.weight"].data_ptr()
        p.memo_conv1_w[d] = sd[m + ".conv1.weight"].data_ptr()
        p.memo_bn1[d] = _bn(sd, m + ".bn1")
        p.memo_conv2_w[d] = sd[m + ".conv2.weight"].data_ptr()
        p.memo_bn2[d] = _bn(sd, m + ".bn2")
        p.memo_conv3_w[d] = sd[m + ".conv3.weight"].data_ptr()
        p.memo_bn3[d] = _bn(sd, m + ".bn3")


def pack_grads(gd, which=3) -> _lib.HeadGrads:
    g = _lib.HeadGrads()
    if which & 1:
        _pack_gce_grads(g, gd)
    if which & 2:
        _pack_trl_grads(g, gd)
    return g


def _pack_gce_grads(g, gd):
    g.glo_fc_w = gd["backbone.glo_fc.0.weight"].data_ptr()
    g.glo_fc_b = gd["backbone.glo_fc.0.bias"].data_ptr()
    g.glo_bn_w = gd["backbone.glo_fc.1.weight"].data_ptr()
    g.glo_bn_b = gd["backbone.glo_fc.1.bias"].data_ptr()
    g.atte0_w = gd["backbone.corr_atte.0.weight"].data_ptr()
    g.atte_bn1_w = gd["backbone.corr_atte.1.weight"].data_ptr()
    g.atte_bn1_b = gd["backbone.corr_atte.1.bias"].data_ptr()
    g.atte2_w = gd["backbone.corr_atte.2.weight"].data_ptr()
    g.atte_bn3_w = gd["backbone.corr_atte.3.weight"].data_ptr()
    g.atte_bn3_b = gd["backbone.corr_atte.3.bias"].data_ptr()
    g.atte5_w = gd["backbone.corr_atte.5.weight"].data_ptr()
    g.atte_bn6_w = gd["backbone.corr_atte.6.weight"].data_ptr()
    g.atte_bn6_b = gd["backbone.corr_atte.6.bias"].data_ptr()


def _pack_trl_grads(g, gd):
    for d, (direction, atte) in enumerate(_DIRS):
        m = TP + "uncorr_memo_" + direction
        g.f1_w[d] = gd[TP + direction + "_f1.0.weight"].data_ptr()
        g.f1_b[d] = gd[TP + direction + "_f1.0.bias"].data_ptr()
        g.f2_w[d] = gd[TP + direction + "_f2.0.weight"].data_ptr()
        g.f2_b[d] = gd[TP + direction + "_f2.0.bias"].data_ptr()
        g.se1_w[d] = gd[TP + "channel_atte_" + atte + "_corr.0.weight"].data_ptr()
        g.se2_w[d] = gd[TP + "channel_atte_" + atte + "_corr.2.weight"].data_ptr()
        g.memo_conv1_w[d] = gd[m + ".conv1.weight"].data_ptr()
        g.memo_bn1_w[d] = gd[m + ".bn1.weight"].data_ptr()
        g.memo_bn1_b[d] = gd[m + ".bn1.bias"].data_ptr()
        g.memo_conv2_w[d] = gd[m + ".conv2.weight"].data_ptr()
        g.memo_bn2_w[d] = gd[m + ".bn2.weight"].data_ptr()
        g.memo_bn2_b[d] = gd[m + ".bn2.bias"].data_ptr()
        g.memo_conv3_w[d] = gd[m + ".conv3.weight"].data_ptr()
        g.memo_bn3_w[d] = gd[m + ".bn3.weight"].data_ptr()
        g.memo_bn3_b[d] = gd[m + ".bn3.bias"].data_ptr()


def workspace_bytes(B, T, save):
    return int(_lib.load_library().grl_head_workspace_bytes(B, T, 1 if save else 0))


def ws_view(ws, B, T, save, name, dtype, shape):
    """Debug/test: typed view of a named intermediate inside the head workspace."""
    off, nbytes = C.c_size_t(), C.c_size_t()
    rc = _lib.load_library().grl_head_ws_lookup(B, T, 1 if save else 0, name.encode(), C.byref(off), C.byref(nbytes))
    if rc != 0:
        raise KeyError(name)
    return ws[off.value:off.value + nbytes.value].view(dtype).view(*shape) if shape else ws[off.value:off.value + nbytes.value].view(dtype)


def saved_state(ws, B, T):
    """Debug/test: every intermediate a train-mode, save_for_backward forward left in the workspace, as fp32 tensors
    in pixel-major layout (names follow DESIGN.md's forward plan; bf16 hi/lo planes are summed)."""
    N, P, R = B * T, B * T * 128, B * 128
    V = lambda name, shape, dt=torch.float32: ws_view(ws, B, T, True, name, dt, shape)
    PL = lambda name, shape: V(name + "_hi", shape, torch.bfloat16).float() + V(name + "_lo", shape, torch.bfloat16).float()
    st = lambda name, c: V(name, (4, c))            # rows: a | c | mean | rstd
    out = dict(X=PL("xp", (P, 2048)), g=V("g", (B, 2048)), u=V("u", (B, 1024)), glo=V("glo", (B, 1024)),
               glo_stat=st("glo_stat", 1024), Y1=PL("y1", (P, 1024)), bn1_stat=st("bn1_stat", 1024),
               Y2=V("y2", (P, 256)), bn2_stat=st("bn2_stat", 256), y3=V("y3", (P,)), bn3_stat=V("bn3_stat", (8,)),
               m=V("m", (P,)), Xc=PL("xc", (P, 2048)), Xu=PL("xu", (P, 2048)), Gc=V("gc", (N, 2048)),
               F2=V("f2", (P, 4096)), mem=PL("mem", (T + 1, 2, R, 2048)), Z=PL("z", (T, 2, R, 2048)),
               F1=V("f1", (T, 2, R, 2048)), q=V("se_q", (T, 2, B, 2048)), h=V("se_h", (T, 2, B, 128)),
               a=V("se_a", (T, 2, B, 2048)), H1=V("h1", (T, 2, R, 512)), H1p=PL("h1p", (T, 2, R, 512)),
               H2=V("h2", (T, 2, R, 512)), H2p=PL("h2p", (T, 2, R, 512)), H3=V("h3", (T, 2, R, 2048)),
               sbn1=V("sbn1", (T, 2, 4, 512)), sbn2=V("sbn2", (T, 2, 4, 512)), sbn3=V("sbn3", (T, 2, 4, 2048)))
    return out


def _alloc_ws(nbytes, device):
    # cudaMalloc'd blocks from the caching allocator are 512-byte aligned; over-allocate to get 1024
    raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    shift = (-raw.data_ptr()) % 1024
    return raw[shift:shift + nbytes]


def head_forward_raw(sd, x, B, T, training, save, want_maps=False, ws=None):
    """One grl_head_forward call.  Returns (f_uncorr, f_corr, corr_map, x_uncorr|None, x_corr|None, workspace)."""
    if not x.is_cuda:
        raise RuntimeError("grl_b200 head needs CUDA tensors (no CPU path exists)")
    x = x.contiguous().float()
    if x.dim() != 4 or x.size(0) != B * T or tuple(x.shape[1:]) != (2048, 16, 8):
        raise RuntimeError("head input must be [b*t, 2048, 16, 8], got %s (b=%d, t=%d)" % (tuple(x.shape), B, T))
    lib = _lib.load_library()
    dev = x.device
    with torch.cuda.device(dev):
        h = _lib.get_handle(dev)
        p = pack_params(sd)
        nbytes = workspace_bytes(B, T, save)
        if ws is None or ws.numel() < nbytes:
            ws = _alloc_ws(nbytes, dev)
        f_uncorr = torch.empty((B, 2048), device=dev)
        f_corr = torch.empty((B, T, 2048), device=dev)
        corr_map = torch.empty((B * T, 1, 16, 8), device=dev)
        xu = torch.empty_like(x) if want_maps else None
        xc = torch.empty_like(x) if want_maps else None
        rc = lib.grl_head_forward(h, C.byref(p), x.data_ptr(), B, T, 1 if training else 0, f_uncorr.data_ptr(), f_corr.data_ptr(),
                                  corr_map.data_ptr(), _lib.ptr(xu), _lib.ptr(xc), ws.data_ptr(), ws.numel(), 1 if save else 0,
                                  _lib.stream_ptr(dev))
        _lib.check(h, rc, "grl_head_forward")
    return f_uncorr, f_corr, corr_map, xu, xc, ws


def head_backward_raw(sd, x, B, T, ws, d_f_uncorr, d_f_corr, d_x_uncorr=None, d_x_corr=None, d_corr_map=None, grads=None):
    """One grl_head_backward call.  Returns (dx, {param name: grad}).  `grads`: optional pre-allocated gradient tensors
    (e.g. views of a flat all-reduce buffer, replicas.GradientAllReduce.views()); they are overwritten."""
    lib = _lib.load_library()
    dev = x.device
    with torch.cuda.device(dev):
        h = _lib.get_handle(dev)
        p = pack_params(sd)
        if grads is None:
            grads = {k: torch.empty_like(sd[k]) for k in head_param_names()}
        g = pack_grads(grads)
        dx = torch.empty_like(x)
        cont = lambda t: None if t is None else t.contiguous().float()
        d_f_uncorr, d_f_corr = cont(d_f_uncorr), cont(d_f_corr)
        d_x_uncorr, d_x_corr, d_corr_map = cont(d_x_uncorr), cont(d_x_corr), cont(d_corr_map)
        rc = lib.grl_head_backward(h, C.byref(p), x.data_ptr(), B, T, d_f_uncorr.data_ptr(), d_f_corr.data_ptr(),
                                   _lib.ptr(d_x_uncorr), _lib.ptr(d_x_corr), _lib.ptr(d_corr_map), dx.data_ptr(), C.byref(g),
                                   ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(h, rc, "grl_head_backward")
    return dx, grads


class GraphedHeadStep(object):
    """One train-mode head step -- grl_head_forward (activations saved) + grl_head_backward -- captured ONCE in a CUDA graph
    and replayed: ~300 kernel launches, the fork/join events of the library's internal side stream and the BatchNorm
    bookkeeping become a single graph launch, so the step costs the host one call and the device no launch gaps.

        step = GraphedHeadStep(sd, B, T)           # sd: {reference state_dict key: CUDA tensor}, parameters AND buffers
        step.x.copy_(layer4_maps); step.d_f_uncorr.copy_(...); step.d_f_corr.copy_(...)
        step()                                     # replay on the current stream
        step.f_uncorr, step.f_corr, step.corr_map, step.dx, step.grads[name]   # static output tensors

    The tensors of `sd` are read (and the BN running buffers / num_batches_tracked updated) in place at every replay, so an
    optimizer that updates the parameters in place composes with it.  Construction leaves `sd` unchanged: the warm-up step
    that precedes the capture runs on zero maps, so every BN buffer it touches is saved first and restored before the capture
    (a loaded checkpoint / resumed run keeps its running statistics).  The reference has no counterpart (it runs eager)."""

    def __init__(self, sd, B, T, device=None):
        dev = torch.device(device) if device is not None else next(iter(sd.values())).device
        self.B, self.T, self.sd = B, T, sd
        self.x = torch.zeros((B * T, 2048, 16, 8), device=dev)
        self.d_f_uncorr = torch.zeros((B, 2048), device=dev)
        self.d_f_corr = torch.zeros((B, T, 2048), device=dev)
        self._ws = _alloc_ws(workspace_bytes(B, T, True), dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        bn_keys = [p_ + s_ for p_ in head_buffer_names() for s_ in (".running_mean", ".running_var", ".num_batches_tracked")
                   if p_ + s_ in sd]
        with torch.cuda.stream(side):                    # warm-up outside the capture: attribute calls, tensor-map cache, allocator
            keep = {k: sd[k].clone() for k in bn_keys}   # the warm-up sees zero maps: its BN statistics must not reach `sd`
            self._run()
            for k in bn_keys:
                sd[k].copy_(keep[k])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._run()

    def _run(self):
        B, T = self.B, self.T
        self.f_uncorr, self.f_corr, self.corr_map, _, _, _ = head_forward_raw(self.sd, self.x, B, T, True, save=True, ws=self._ws)
        self.dx, self.grads = head_backward_raw(self.sd, self.x, B, T, self._ws, self.d_f_uncorr, self.d_f_corr)
        with torch.no_grad():                            # torch/nn/modules/batchnorm.py: +1 per BN call in train mode
            keys = [p_ + ".num_batches_tracked" for p_ in head_buffer_names() if p_ + ".num_batches_tracked" in self.sd]
            per_step = [self.sd[k] for k in keys if "uncorr_memo" in k]
            once = [self.sd[k] for k in keys if "uncorr_memo" not in k]
            if per_step:
                torch._foreach_add_(per_step, T)
            if once:
                torch._foreach_add_(once, 1)

    def __call__(self):
        self.graph.replay()
        return self.f_uncorr, self.f_corr, self.dx, self.grads


def _check_maps(t, B, T, what):
    if not t.is_cuda:
        raise RuntimeError("grl_b200 head needs CUDA tensors (no CPU path exists)")
    t = t.contiguous().float()
    if t.numel() != B * T * 2048 * 128 or tuple(t.shape[-3:]) != (2048, 16, 8):
        raise RuntimeError("%s must be [b*t, 2048, 16, 8], got %s (b=%d, t=%d)" % (what, tuple(t.shape), B, T))
    return t


def gce_forward_raw(sd, x, B, T, training, save, ws=None):
    """grl_gce_forward: (x_uncorr, x_corr, corr_map, workspace)."""
    x = _check_maps(x, B, T, "GCE input")
    lib = _lib.load_library()
    dev = x.device
    with torch.cuda.device(dev):
        h = _lib.get_handle(dev)
        p = pack_params(sd, 1)
        nbytes = workspace_bytes(B, T, save)
        if ws is None or ws.numel() < nbytes:
            ws = _alloc_ws(nbytes, dev)
        xu, xc = torch.empty_like(x), torch.empty_like(x)
        corr_map = torch.empty((B * T, 1, 16, 8), device=dev)
        rc = lib.grl_gce_forward(h, C.byref(p), x.data_ptr(), B, T, 1 if training else 0, xu.data_ptr(), xc.data_ptr(),
                                 corr_map.data_ptr(), ws.data_ptr(), ws.numel(), 1 if save else 0, _lib.stream_ptr(dev))
        _lib.check(h, rc, "grl_gce_forward")
    return xu, xc, corr_map, ws


def gce_backward_raw(sd, B, T, ws, d_x_uncorr, d_x_corr, d_corr_map):
    """grl_gce_backward: (dx, {GCE param name: grad})."""
    lib = _lib.load_library()
    dev = ws.device
    with torch.cuda.device(dev):
        h = _lib.get_handle(dev)
        p = pack_params(sd, 1)
        grads = {k: torch.empty_like(sd[k]) for k in gce_param_names()}
        g = pack_grads(grads, 1)
        dx = torch.empty((B * T, 2048, 16, 8), device=dev)
        cont = lambda t: None if t is None else t.contiguous().float()
        d_x_uncorr, d_x_corr, d_corr_map = cont(d_x_uncorr), cont(d_x_corr), cont(d_corr_map)
        rc = lib.grl_gce_backward(h, C.byref(p), B, T, _lib.ptr(d_x_uncorr), _lib.ptr(d_x_corr), _lib.ptr(d_corr_map),
                                  dx.data_ptr(), C.byref(g), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(h, rc, "grl_gce_backward")
    return dx, grads


def trl_forward_raw(sd, x_uncorr, x_corr, B, T, training, save, ws=None):
    """grl_trl_forward: (f_uncorr, f_corr, workspace)."""
    x_uncorr = _check_maps(x_uncorr, B, T, "x_uncorr")
    x_corr = _check_maps(x_corr, B, T, "x_corr")
    lib = _lib.load_library()
    dev = x_corr.device
    with torch.cuda.device(dev):
        h = _lib.get_handle(dev)
        p = pack_params(sd, 2)
        nbytes = workspace_bytes(B, T, save)
        if ws is None or ws.numel() < nbytes:
            ws = _alloc_ws(nbytes, dev)
        f_uncorr = torch.empty((B, 2048), device=dev)
        f_corr = torch.empty((B, T, 2048), device=dev)
        rc = lib.grl_trl_forward(h, C.byref(p), x_uncorr.data_ptr(), x_corr.data_ptr(), B, T, 1 if training else 0,
                                 f_uncorr.data_ptr(), f_corr.data_ptr(), ws.data_ptr(), ws.numel(), 1 if save else 0,
                                 _lib.stream_ptr(dev))
        _lib.check(h, rc, "grl_trl_forward")
    return f_uncorr, f_corr, ws


def trl_backward_raw(sd, B, T, ws, d_f_uncorr, d_f_corr):
    """grl_trl_backward: (d_x_uncorr, d_x_corr [b*t,2048,16,8], {TRL param name: grad})."""
    lib = _lib.load_library()
    dev = ws.device
    with torch.cuda.device(dev):
        h = _lib.get_handle(dev)
        p = pack_params(sd, 2)
        grads = {k: torch.empty_like(sd[k]) for k in trl_param_names()}
        g = pack_grads(grads, 2)
        dxu = torch.empty((B * T, 2048, 16, 8), device=dev)
        dxc = torch.empty((B * T, 2048, 16, 8), device=dev)
        d_f_uncorr, d_f_corr = d_f_uncorr.contiguous().float(), d_f_corr.contiguous().float()
        rc = lib.grl_trl_backward(h, C.byref(p), B, T, d_f_uncorr.data_ptr(), d_f_corr.data_ptr(), dxu.data_ptr(), dxc.data_ptr(),
                                  C.byref(g), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(h, rc, "grl_trl_backward")
    return dxu, dxc, grads


class _GceFunction(torch.autograd.Function):
    """Stand-alone GCE: (x, *gce params) -> (x_uncorr, x_corr, corr_map)."""

    @staticmethod
    def forward(ctx, x, B, T, training, need_grad, names, sd_buffers, *params):
        sd = dict(zip(names, params))
        sd.update(sd_buffers)
        if need_grad and not training:
            raise RuntimeError("grl_b200 GCE: backward with eval-mode BatchNorm is not supported")
        xu, xc, corr_map, ws = gce_forward_raw(sd, x, B, T, training, save=need_grad)
        ctx.B, ctx.T, ctx.names, ctx.sd_buffers = B, T, names, sd_buffers
        ctx.ws = ws if need_grad else None
        ctx.save_for_backward(*params)
        return xu, xc, corr_map

    @staticmethod
    def backward(ctx, g_xu, g_xc, g_map):
        sd = dict(zip(ctx.names, ctx.saved_tensors))
        sd.update(ctx.sd_buffers)
        dx, grads = gce_backward_raw(sd, ctx.B, ctx.T, ctx.ws, g_xu, g_xc, g_map)
        ctx.ws = None
        return (dx, None, None, None, None, None, None) + tuple(grads[k] for k in ctx.names)


class _TrlFunction(torch.autograd.Function):
    """Stand-alone TRL: (x_uncorr, x_corr, *trl params) -> (f_uncorr, f_corr)."""

    @staticmethod
    def forward(ctx, x_uncorr, x_corr, B, T, training, need_grad, names, sd_buffers, *params):
        sd = dict(zip(names, params))
        sd.update(sd_buffers)
        if need_grad and not training:
            raise RuntimeError("grl_b200 TRL: backward with eval-mode BatchNorm is not supported")
        f_uncorr, f_corr, ws = trl_forward_raw(sd, x_uncorr, x_corr, B, T, training, save=need_grad)
        ctx.B, ctx.T, ctx.names, ctx.sd_buffers, ctx.shape = B, T, names, sd_buffers, x_corr.shape
        ctx.ws = ws if need_grad else None
        ctx.save_for_backward(*params)
        return f_uncorr, f_corr

    @staticmethod
    def backward(ctx, g_fu, g_fc):
        sd = dict(zip(ctx.names, ctx.saved_tensors))
        sd.update(ctx.sd_buffers)
        dev = ctx.ws.device
        g_fu = torch.zeros((ctx.B, 2048), device=dev) if g_fu is None else g_fu
        g_fc = torch.zeros((ctx.B, ctx.T, 2048), device=dev) if g_fc is None else g_fc
        dxu, dxc, grads = trl_backward_raw(sd, ctx.B, ctx.T, ctx.ws, g_fu, g_fc)
        ctx.ws = None
        return (dxu.view(ctx.shape), dxc.view(ctx.shape), None, None, None, None, None, None) + tuple(grads[k] for k in ctx.names)


def _bump_bn(buffers, prefixes, steps_of):
    with torch.no_grad():
        for prefix in prefixes:
            buffers[prefix + ".num_batches_tracked"] += steps_of(prefix)


class _HeadFunction(torch.autograd.Function):
    """autograd node for the fused head: (x, *params) -> (f_uncorr, f_corr, corr_map, x_uncorr, x_corr)."""

    @staticmethod
    def forward(ctx, x, B, T, training, want_maps, need_grad, names, sd_buffers, *params):
        sd = dict(zip(names, params))
        sd.update(sd_buffers)
        if need_grad and not training:
            raise RuntimeError("grl_b200 head: backward with eval-mode BatchNorm is not supported; call .train() "
                               "or wrap inference in torch.no_grad()")
        f_uncorr, f_corr, corr_map, xu, xc, ws = head_forward_raw(sd, x, B, T, training, save=need_grad, want_maps=want_maps)
        ctx.B, ctx.T, ctx.names, ctx.sd_buffers, ctx.want_maps = B, T, names, sd_buffers, want_maps
        ctx.ws = ws if need_grad else None
        ctx.save_for_backward(x, *params)
        if not want_maps:
            xu = x.new_empty(0)
            xc = x.new_empty(0)
        if not want_maps:
            ctx.mark_non_differentiable(xu, xc)
        return f_uncorr, f_corr, corr_map, xu, xc

    @staticmethod
    def backward(ctx, g_fu, g_fc, g_map, g_xu, g_xc):
        x, *params = ctx.saved_tensors
        sd = dict(zip(ctx.names, params))
        sd.update(ctx.sd_buffers)
        B, T = ctx.B, ctx.T
        zeros = lambda shape: torch.zeros(shape, device=x.device)
        g_fu = zeros((B, 2048)) if g_fu is None else g_fu
        g_fc = zeros((B, T, 2048)) if g_fc is None else g_fc
        if not ctx.want_maps:
            g_xu = g_xc = None
        dx, grads = head_backward_raw(sd, x.contiguous().float(), B, T, ctx.ws, g_fu, g_fc, g_xu, g_xc, g_map)
        ctx.ws = None
        return (dx, None, None, None, None, None, None, None) + tuple(grads[k] for k in ctx.names)


def _collect(module_sd_items, prefix_map):
    out = {}
    for name, t in module_sd_items:
        out[prefix_map + name] = t
    return out


class _HeadState:
    """Gathers parameters/buffers of a Backbone (GCE part) and a TRLBlock under the reference's key names."""

    @staticmethod
    def tensors(backbone, trl):
        params, buffers = {}, {}
        for n, t in backbone.named_parameters():
            if not n.startswith("base."):
                params["backbone." + n] = t
        for n, t in backbone.named_buffers():
            if not n.startswith("base."):
                buffers["backbone." + n] = t
        for n, t in trl.named_parameters():
            params[TP + n] = t
        for n, t in trl.named_buffers():
            buffers[TP + n] = t
        return params, buffers


def run_head(backbone, trl, x, b, t, want_maps=False):
    """Fused GCE+TRL through the C ABI with autograd.  x: layer4 maps [b*t, 2048, 16, 8]."""
    params, buffers = _HeadState.tensors(backbone, trl)
    names = head_param_names()
    training = backbone.training
    plist = [params[k] for k in names]
    # grad mode is off inside Function.forward, so decide here whether activations must be kept for a backward
    need_grad = torch.is_grad_enabled() and (x.requires_grad or any(t_.requires_grad for t_ in plist))
    out = _HeadFunction.apply(x, b, t, training, want_maps, need_grad, names, buffers, *plist)
    if training:    # num_batches_tracked bookkeeping (torch/nn/modules/batchnorm.py): +1 per BN call; two fused launches
        with torch.no_grad():
            per_step = [buffers[p_ + ".num_batches_tracked"] for p_ in head_buffer_names() if "uncorr_memo" in p_]
            once = [buffers[p_ + ".num_batches_tracked"] for p_ in head_buffer_names() if "uncorr_memo" not in p_]
            if per_step:
                torch._foreach_add_(per_step, t)
            if once:
                torch._foreach_add_(once, 1)
    return out


# --------------------------------------------------------------------------------------------
# Drop-in modules (same attribute names as the reference => identical state_dict keys)
# --------------------------------------------------------------------------------------------
def _resnet50_stride1_base():
    """conv1..layer4 of a torchvision-style ResNet-50 with layer4 stride 1 (reid/models/resnets1.py:109).
    Stays in PyTorch/cuDNN by design (north-star); weights are loaded by the caller via state_dict."""
    import torchvision
    r = torchvision.models.resnet50(weights=None)
    r.layer4[0].conv2.stride = (1, 1)
    r.layer4[0].downsample[0].stride = (1, 1)
    return nn.Sequential(r.conv1, r.bn1, nn.ReLU(), r.maxpool, r.layer1, r.layer2, r.layer3, r.layer4)


class Backbone(nn.Module):
    """reid/models/basebranch.py:21-68.  `base` is cuDNN ResNet-50; GCE runs in libgrl_b200."""

    def __init__(self, height=256, width=128, base=None):
        super(Backbone, self).__init__()
        self.base = _resnet50_stride1_base() if base is None else base
        self.glo_fc = nn.Sequential(nn.Linear(2048, 1024), nn.BatchNorm1d(1024), nn.ReLU())
        self.corr_atte = nn.Sequential(
            nn.Conv2d(2048 + 1024, 1024, 1, 1, bias=False), nn.BatchNorm2d(1024),
            nn.Conv2d(1024, 256, 1, 1, bias=False), nn.BatchNorm2d(256), nn.ReLU(),
            nn.Conv2d(256, 1, 1, 1, bias=False), nn.BatchNorm2d(1))
        self._trl_for_fused = None      # set by ResNet50_GRL_Model (not a submodule: no state_dict change)

    def gce(self, x, b, t):
        """GCE on layer4 maps x [b*t, 2048, 16, 8] (grl_gce_forward / grl_gce_backward)."""
        params = {"backbone." + n: p for n, p in self.named_parameters() if not n.startswith("base.")}
        buffers = {"backbone." + n: p for n, p in self.named_buffers() if not n.startswith("base.")}
        names = gce_param_names()
        plist = [params[k] for k in names]
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in plist))
        out = _GceFunction.apply(x, b, t, self.training, need_grad, names, buffers, *plist)
        if self.training:
            _bump_bn(buffers, [k for k in head_buffer_names() if k.startswith("backbone.")], lambda _: 1)
        return out

    def forward(self, x, b, t):
        """reid/models/basebranch.py:52-68: returns (x_uncorr, x_corr, corr_map)."""
        x = self.base(x)
        return self.gce(x, b, t)


class BasicBlock(nn.Module):
    """reid/models/grl_model.py:51-85.  Inside TRLBlock the math runs in the fused recurrence kernels (forward and backward);
    called on its own, forward(x1, x2) runs one memory update through grl_basic_block_forward (forward only)."""

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super(BasicBlock, self).__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU()

    def forward(self, x1, x2):
        """grl_model.py:67-85: relu(bn3(conv3(relu(bn2(conv2(relu(bn1(conv1(x1 + x2)))))))) + (x1 + x2)); x* [n, 2048, 16, 8]."""
        if torch.is_grad_enabled() and (x1.requires_grad or x2.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("grl_b200 BasicBlock.forward is forward-only (call it under torch.no_grad()); its gradients are "
                               "computed inside TRLBlock's fused backward")
        if not (x1.is_cuda and x2.is_cuda):
            raise RuntimeError("grl_b200 head needs CUDA tensors (no CPU path exists)")
        x1, x2 = x1.contiguous().float(), x2.contiguous().float()
        if x1.shape != x2.shape or x1.dim() != 4 or tuple(x1.shape[1:]) != (2048, 16, 8) or self.conv1.weight.shape != (512, 2048, 1, 1):
            raise RuntimeError("BasicBlock.forward: the memory block works on [n, 2048, 16, 8] maps, got %s / %s" % (tuple(x1.shape), tuple(x2.shape)))
        lib = _lib.load_library()
        dev, n = x1.device, x1.size(0)
        bn = []
        for m in (self.bn1, self.bn2, self.bn3):
            r = _lib.BnParams()
            r.weight, r.bias = m.weight.data_ptr(), m.bias.data_ptr()
            r.running_mean, r.running_var = m.running_mean.data_ptr(), m.running_var.data_ptr()
            bn.append(r)
        with torch.cuda.device(dev):
            h = _lib.get_handle(dev)
            out = torch.empty_like(x1)
            ws = _alloc_ws(int(lib.grl_basic_block_workspace_bytes(n)), dev)
            rc = lib.grl_basic_block_forward(h, self.conv1.weight.data_ptr(), C.byref(bn[0]), self.conv2.weight.data_ptr(), C.byref(bn[1]),
                                             self.conv3.weight.data_ptr(), C.byref(bn[2]), x1.data_ptr(), x2.data_ptr(), n,
                                             1 if self.training else 0, out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
            _lib.check(h, rc, "grl_basic_block_forward")
        if self.training:
            with torch.no_grad():
                for m in (self.bn1, self.bn2, self.bn3):
                    m.num_batches_tracked += 1
        return out


class TRLBlock(nn.Module):
    """reid/models/grl_model.py:87-180 (note the reference's spelling `channel_atte_foreward_corr`)."""

    def __init__(self, feat_num):
        super(TRLBlock, self).__init__()
        self.feat_num = feat_num
        self.feat_num_half = int(feat_num / 2)
        self.uncorr_memo_forward = BasicBlock(2048, 512)
        self.forward_f1 = nn.Sequential(nn.Conv2d(2048, 2048, 1, 1), nn.ReLU())
        self.forward_f2 = nn.Sequential(nn.Conv2d(2048, 2048, 1, 1), nn.ReLU())
        self.channel_atte_foreward_corr = nn.Sequential(nn.Linear(2048, 2048 // 16, bias=False), nn.ReLU(inplace=True),
                                                        nn.Linear(2048 // 16, 2048, bias=False), nn.Sigmoid())
        self.uncorr_memo_backward = BasicBlock(2048, 512)
        self.backward_f1 = nn.Sequential(nn.Conv2d(2048, 2048, 1, 1), nn.ReLU())
        self.backward_f2 = nn.Sequential(nn.Conv2d(2048, 2048, 1, 1), nn.ReLU())
        self.channel_atte_backward_corr = nn.Sequential(nn.Linear(2048, 2048 // 16, bias=False), nn.ReLU(inplace=True),
                                                        nn.Linear(2048 // 16, 2048, bias=False), nn.Sigmoid())

    def forward(self, x_uncorr, x_corr):
        """reid/models/grl_model.py:131-180: x_* [b, t, 2048, 16, 8] -> (f_uncorr [b,2048], f_corr [b,t,2048])."""
        b, t = x_corr.size(0), x_corr.size(1)
        params = {TP + n: p for n, p in self.named_parameters()}
        buffers = {TP + n: p for n, p in self.named_buffers()}
        names = trl_param_names()
        plist = [params[k] for k in names]
        need_grad = torch.is_grad_enabled() and (x_uncorr.requires_grad or x_corr.requires_grad or any(p.requires_grad for p in plist))
        out = _TrlFunction.apply(x_uncorr, x_corr, b, t, self.training, need_grad, names, buffers, *plist)
        if self.training:
            _bump_bn(buffers, [k for k in head_buffer_names() if k.startswith(TP)], lambda _: t)
        return out


class ResNet50_GRL_Model(nn.Module):
    """reid/models/grl_model.py:184-228.  forward(inputs[B,T,3,H,W]) -> (x_uncorr [B,2048], x_corr [B,T,2048])."""

    def __init__(self, num_feat=2048, num_features=512, height=256, width=128, pretrained=True, dropout=0, numclasses=0,
                 base=None):
        super(ResNet50_GRL_Model, self).__init__()
        self.pretrained = pretrained
        self.num_feat = num_feat
        self.dropout = dropout
        self.num_classes = numclasses
        self.output_dim = num_features
        self.backbone = Backbone(height=height, width=width, base=base)
        self.temporal_learning_block = TRLBlock(2048)
        self.corr_bn = nn.BatchNorm1d(2048)
        nn.init.constant_(self.corr_bn.weight, 1)
        nn.init.constant_(self.corr_bn.bias, 0)
        self.uncorr_bn = nn.BatchNorm1d(2048)
        nn.init.constant_(self.uncorr_bn.weight, 1)
        nn.init.constant_(self.uncorr_bn.bias, 0)

    def head(self, feat, b, t, want_maps=False):
        """GCE + TRL on layer4 maps feat [b*t, 2048, 16, 8] (fused C-ABI call)."""
        return run_head(self.backbone, self.temporal_learning_block, feat, b, t, want_maps)

    def forward(self, inputs, training=True):
        b, t, c, h, w = inputs.size()
        im_input = inputs.view(b * t, c, h, w)
        feat = self.backbone.base(im_input)                                  # PyTorch / cuDNN
        f_uncorr, f_corr, _, _, _ = self.head(feat, b, t)                    # libgrl_b200 (sm_100a)
        x_corr = self.corr_bn(f_corr.view(b * t, 2048)).view(b, t, 2048)     # grl_model.py:222-226 (tiny tail, PyTorch)
        x_corr = F.normalize(x_corr, p=2, dim=2)
        x_uncorr = self.uncorr_bn(f_uncorr.view(b, 2048)).view(b, 2048)
        x_uncorr = F.normalize(x_uncorr, p=2, dim=1)
        return x_uncorr, x_corr


def resnet50_grl(*args, **kwargs):
    return ResNet50_GRL_Model(*args, **kwargs)
