"""ORACLE — test infrastructure, not product code.

CPU restatement (plain PyTorch fp32 functional ops; none of the reference's nn.Modules, no CUDA extension) of the
reference's TDN ResNet-50 path (BASELINE configs[4]): workoutdetector/models/tdn.py +
workoutdetector/models/tsn.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu arms may import this file.

Pinned against the reference itself: oracle/gen_golden.py builds the reference's ``tdn.create_model`` (unmodified,
from /root/reference, behind import shims), loads the seeded state_dict made by ``random_state_dict`` below with
``load_state_dict(strict=True)`` — which also pins every key name and shape — and asserts that ``tdn_forward``
reproduces the module's logits and hooked activations; tests/golden/tdn_golden.npz keeps the reference outputs.
"""
import math
from collections import OrderedDict
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

BLOCKS = (3, 4, 6, 3)
PLANES = (64, 128, 256, 512)
BN_EPS = 1e-5


# --------------------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------------------
def random_state_dict(num_class: int, seed: int, num_segments: int = 8) -> "OrderedDict[str, torch.Tensor]":
    """A seeded state_dict with exactly the keys / shapes of the reference's TSN(TDN_Net) module (tdn.py:92-137,
    188-249, 339-364, 475-495, 523-567; tsn.py:137-169) and non-trivial values everywhere: He-scaled convolutions
    WITH biases (the FBResNet convs have bias=True), BatchNorm statistics / affine terms away from identity, bn3 gamma
    in [0.3, 0.6] so the residual stream stays O(1), temporal-shift Conv1d kernels = the reference's 'shift' init
    (tdn.py:352-358) plus a perturbation so all three taps matter."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()

    def conv(name, cout, cin, k, bias):
        sd[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (cin * k * k))
        if bias:
            sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.05

    def bn(name, n, lo=0.8, hi=1.2):
        sd[name + ".weight"] = torch.rand(n, generator=g) * (hi - lo) + lo
        sd[name + ".bias"] = torch.randn(n, generator=g) * 0.05
        sd[name + ".running_mean"] = torch.randn(n, generator=g) * 0.1
        sd[name + ".running_var"] = torch.rand(n, generator=g) + 0.5
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    def bottleneck(p, inplanes, planes, down, shift):
        conv(p + ".conv1", planes, inplanes, 1, True)
        bn(p + ".bn1", planes)
        if shift:  # BottleneckShift registers mse and shift between bn1 and conv2 (tdn.py:475-484)
            r = planes // 16
            m = p + ".mse"
            conv(m + ".conv1", r, planes, 1, False)
            bn(m + ".bn1", r)
            sd[m + ".conv2.weight"] = torch.randn(r, 1, 3, 3, generator=g) * 0.3
            conv(m + ".conv3", planes, r, 1, False)
            sd[m + ".conv3.weight"] *= 2.0
            bn(m + ".bn3", planes)
            conv(m + ".conv3_smallscale2", r, r, 3, False)
            bn(m + ".bn3_smallscale2", r)
            conv(m + ".conv3_smallscale4", r, r, 3, False)
            bn(m + ".bn3_smallscale4", r)
            fold = planes // 8
            w = torch.zeros(planes, 1, 3)
            w[:fold, 0, 2] = 1
            w[fold:2 * fold, 0, 0] = 1
            w[2 * fold:, 0, 1] = 1
            sd[p + ".shift.conv.weight"] = w + torch.randn(planes, 1, 3, generator=g) * 0.1
        conv(p + ".conv2", planes, planes, 3, True)
        bn(p + ".bn2", planes)
        conv(p + ".conv3", planes * 4, planes, 1, True)
        bn(p + ".bn3", planes * 4, 0.3, 0.6)
        if down:
            conv(p + ".downsample.0", planes * 4, inplanes, 1, True)
            bn(p + ".downsample.1", planes * 4, 0.3, 0.6)

    def layer(prefix, L, shift):
        inplanes = 64 if L == 0 else PLANES[L - 1] * 4
        for b in range(BLOCKS[L]):
            bottleneck(f"{prefix}.{b}", inplanes, PLANES[L], b == 0, shift)
            inplanes = PLANES[L] * 4

    # registration order of TDN_Net.__init__ (tdn.py:100-135)
    conv("base_model.conv1", 64, 3, 7, True)
    bn("base_model.bn1", 64)
    conv("base_model.conv1_temp", 64, 3, 7, True)   # kept as a module attribute, unused by forward (tdn.py:105)
    conv("base_model.conv1_5.0", 64, 12, 7, False)
    bn("base_model.conv1_5.1", 64)
    layer("base_model.resnext_layer1", 0, False)
    layer("base_model.layer1_bak", 0, False)
    layer("base_model.layer2_bak", 1, True)
    layer("base_model.layer3_bak", 2, True)
    layer("base_model.layer4_bak", 3, True)
    sd["new_fc.weight"] = torch.randn(num_class, 2048, generator=g) * 0.05
    sd["new_fc.bias"] = torch.randn(num_class, generator=g) * 0.1
    return sd


def golden_input(seed: int = 2) -> torch.Tensor:
    """The seeded [2, 8, 5, 3, 224, 224] input of tests/golden/tdn_golden.npz: clip 0 is white noise, clip 1 is one
    random image plus 15 % noise per frame (small frame differences, what real video looks like)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(2, 8, 5, 3, 224, 224, generator=g)
    base = torch.randn(1, 1, 1, 3, 224, 224, generator=g)
    x[1] = (base + 0.15 * torch.randn(1, 8, 5, 3, 224, 224, generator=g))[0]
    return x


# --------------------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------------------
def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, BN_EPS)


def _r(x, emulate):
    return x.to(torch.bfloat16).to(torch.float32) if emulate else x


def mse_module(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor, n_segment: int = 8) -> torch.Tensor:
    """mSEModule.forward (tdn.py:266-334): motion-excitation gate, ``x + x * y``."""
    nt, c, h, w = x.shape
    r = c // 16
    bott = _bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"]))                       # 267-268
    rb = bott.view(-1, n_segment, r, h, w)
    cb = F.conv2d(bott, sd[p + ".conv2.weight"], padding=1, groups=r).view(-1, n_segment, r, h, w)   # 277
    d_f = cb[:, 1:] - rb[:, :-1]                                                           # 284
    d_b = cb[:, :-1] - rb[:, 1:]                                                           # 285
    zero = torch.zeros_like(rb[:, :1])
    d_f = torch.cat([d_f, zero], dim=1).reshape(nt, r, h, w)                               # pad (…,0,1): zero at t = T-1
    d_b = torch.cat([zero, d_b], dim=1).reshape(nt, r, h, w)                               # pad (…,1,0): zero at t = 0

    def branch(d):
        s2 = _bn(sd, p + ".bn3_smallscale2", F.conv2d(F.avg_pool2d(d, 2, 2), sd[p + ".conv3_smallscale2.weight"],
                                                      padding=1))                           # 299-308
        s4 = _bn(sd, p + ".bn3_smallscale4", F.conv2d(d, sd[p + ".conv3_smallscale4.weight"], padding=1))  # 310-313
        s2 = F.interpolate(s2, d.shape[2:])                                                # nearest, 315-318
        y = _bn(sd, p + ".bn3", F.conv2d(1.0 / 3.0 * d + 1.0 / 3.0 * s2 + 1.0 / 3.0 * s4, sd[p + ".conv3.weight"]))
        return torch.sigmoid(y) - 0.5                                                      # 320-330

    y = 0.5 * branch(d_f) + 0.5 * branch(d_b)                                              # 332
    return x + x * y                                                                       # 333


def shift_module(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor, n_segment: int = 8) -> torch.Tensor:
    """ShiftModule.forward (tdn.py:366-376): depthwise Conv1d (k=3, pad=1, no bias) along the segment axis."""
    nt, c, h, w = x.shape
    n = nt // n_segment
    y = x.view(n, n_segment, c, h, w).permute(0, 3, 4, 2, 1).reshape(n * h * w, c, n_segment)
    y = F.conv1d(y, sd[p + ".conv.weight"], padding=1, groups=c)
    return y.view(n, h, w, c, n_segment).permute(0, 4, 3, 1, 2).reshape(nt, c, h, w)


def diff_input(x: torch.Tensor) -> torch.Tensor:
    """tdn.py:147-150: the four frame differences of a 5-frame segment, average-pooled 2x2.
    x [N*T, 15, H, W] -> [N*T, 12, H/2, W/2]."""
    t = x[:, 3:15] - x[:, 0:12]
    return F.avg_pool2d(t, 2, 2)


def tdn_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, num_segments: int = 8, alpha: float = 0.5,
                beta: float = 0.5, emulate_bf16: bool = False,
                tap: Optional[Callable[[str, torch.Tensor], None]] = None) -> torch.Tensor:
    """TSN.forward (tsn.py:335-351) around TDN_Net.forward (tdn.py:139-178): x [B,T,5,3,H,W] or [B*T*5,3,H,W] or
    [B*T,15,H,W] fp32 (normalised) -> consensus logits [B, num_class].

    emulate_bf16=True restates what the bf16 engine computes: BN and conv bias folded into bf16 weights / fp32
    bias, every stored activation rounded to bf16, the motion-excitation arithmetic in fp32 on those activations."""
    sd = {k: v.to(torch.float32) for k, v in sd.items() if v.is_floating_point()}
    emu = emulate_bf16
    x = x.to(torch.float32).reshape((-1, 15) + tuple(x.shape[-2:]))                        # tsn.py:337-338

    def emit(name, t):
        if tap is not None:
            tap(name, t)

    def conv_bn(x, cv, bn, stride, pad, relu, residual=None):
        w = sd[cv + ".weight"]
        b = sd.get(cv + ".bias")
        if emu:
            s = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + BN_EPS)
            bias = sd[bn + ".bias"] - sd[bn + ".running_mean"] * s + (b * s if b is not None else 0.0)
            y = F.conv2d(x, _r(w * s.view(-1, 1, 1, 1), True), bias, stride=stride, padding=pad)
        else:
            y = _bn(sd, bn, F.conv2d(x, w, b, stride=stride, padding=pad))
        if residual is not None:
            y = y + residual
        if relu:
            y = F.relu(y)
        return _r(y, emu)

    def bottleneck(x, p, n, stride, shift):
        y = conv_bn(x, p + ".conv1", p + ".bn1", 1, 0, True)
        emit(n + ".conv1", y)
        if shift:                                                                          # tdn.py:503-504
            y = _r(shift_module(sd, p + ".shift", mse_module(sd, p + ".mse", y, num_segments), num_segments), emu)
            emit(n + ".mse", y)
        y = conv_bn(y, p + ".conv2", p + ".bn2", stride, 1, True)
        emit(n + ".conv2", y)
        if (p + ".downsample.0.weight") in sd:
            idt = conv_bn(x, p + ".downsample.0", p + ".downsample.1", stride, 0, False)
            emit(n + ".downsample", idt)
        else:
            idt = x
        y = conv_bn(y, p + ".conv3", p + ".bn3", 1, 0, True, residual=idt)
        emit(n + ".conv3", y)
        return y

    def layer(x, prefix, name, L, shift):
        for b in range(BLOCKS[L]):
            x = bottleneck(x, f"base_model.{prefix}.{b}", f"{name}.{b}", 2 if (L > 0 and b == 0) else 1, shift)
        return x

    d = _r(diff_input(x), emu)                                                             # tdn.py:147-150
    emit("diff_in", d)
    xd = conv_bn(d, "base_model.conv1_5.0", "base_model.conv1_5.1", 2, 3, True)            # 150
    emit("conv1_5", xd)
    xd = F.max_pool2d(xd, 3, 2, 1)                                                         # 152
    emit("maxpool_diff", xd)
    t1 = xd
    xd = layer(xd, "resnext_layer1", "diff1", 0, False)                                    # 155
    xc = conv_bn(_r(x[:, 6:9], emu), "base_model.conv1", "base_model.bn1", 2, 3, True)     # 157-159, centre frame
    emit("conv1", xc)
    xc = F.max_pool2d(xc, 3, 2, 1)                                                         # 161
    emit("maxpool", xc)
    xc = _r(alpha * xc + beta * F.interpolate(t1, xc.shape[2:]), emu)                      # 162-163
    emit("fuse1", xc)
    xc = layer(xc, "layer1_bak", "layer1", 0, False)                                       # 165
    xc = _r(alpha * xc + beta * F.interpolate(xd, xc.shape[2:]), emu)                      # 166-167
    emit("fuse2", xc)
    xc = layer(xc, "layer2_bak", "layer2", 1, True)                                        # 169-171
    xc = layer(xc, "layer3_bak", "layer3", 2, True)
    xc = layer(xc, "layer4_bak", "layer4", 3, True)
    f = xc.mean(dim=(2, 3))                                                                # avgpool, 173-174
    o = F.linear(f, sd["new_fc.weight"], sd["new_fc.bias"])                                # fc = Dropout (eval: id), tsn.py:343-344
    return o.view(-1, num_segments, o.shape[1]).mean(dim=1)                                # tsn.py:349-351
