"""ORACLE — test infrastructure, not product code.

CPU restatement (plain PyTorch fp32 functional ops, no nn.Module of the reference, no CUDA extension) of the
reference's TSM ResNet-50 hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this file.  The product (workoutdetector_b200/) never does.

Pinned against the reference itself: oracle/gen_golden.py imports /root/reference/workoutdetector (with import
shims) in the build container, checks every function here against the reference module on seeded inputs and writes
tests/golden/*.  tests/test_oracle_golden.py re-checks this file against those committed fixtures everywhere.

Each function cites the reference lines it restates.
"""
from collections import OrderedDict
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

BLOCKS = (3, 4, 6, 3)
PLANES = (64, 128, 256, 512)
BN_EPS = 1e-5
MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


# --------------------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------------------
def reference_init_state_dict(num_class: int, seed: int, num_segments: int = 8) -> "OrderedDict[str, torch.Tensor]":
    """The state_dict the reference's ``create_model(num_class, 8, 'resnet50')`` has after ``torch.manual_seed(seed)``
    when no pretrained weights are available (workoutdetector/models/tsm.py:212-262):
    torchvision resnet50 init, then fc = Linear(2048, C) drawn twice (tsm.py:246-248 and 259-262), N(0, 0.001) / 0.
    Key names carry the TemporalShift ``.net`` wrapper of every bottleneck conv1 (tsm.py:125-137)."""
    import torchvision

    torch.manual_seed(seed)
    net = torchvision.models.resnet50(weights=None)
    fc = torch.nn.Linear(2048, num_class)
    torch.nn.init.normal_(fc.weight, 0, 0.001)
    torch.nn.init.constant_(fc.bias, 0)
    fc = torch.nn.Linear(2048, num_class)
    torch.nn.init.normal_(fc.weight, 0, 0.001)
    torch.nn.init.constant_(fc.bias, 0)
    sd = OrderedDict()
    for k, v in net.state_dict().items():
        if k.startswith("fc.") or k.endswith("num_batches_tracked"):
            continue
        parts = k.split(".")
        if len(parts) >= 3 and parts[0].startswith("layer") and parts[2] == "conv1":
            k = ".".join(parts[:3] + ["net"] + parts[3:])
        sd["base_model." + k] = v.clone()
    sd["fc.weight"] = fc.weight.detach().clone()
    sd["fc.bias"] = fc.bias.detach().clone()
    return sd


def randomize_bn_and_fc(sd: Dict[str, torch.Tensor], seed: int, fc_std: float = 0.05) -> Dict[str, torch.Tensor]:
    """Test-only weight variant: non-trivial BatchNorm statistics / affine terms (so BN folding is exercised). bn3
    gamma in [0.3, 0.6] keeps the residual stream bounded (|x| stays O(1) instead of the hundreds the identity-BN
    init reaches) while leaving ~6 % clip-to-clip variation in the pooled features; fc is widened so logits move.
    Deterministic in ``seed``; the reference module loads it with load_state_dict like any checkpoint."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict((k, v.clone()) for k, v in sd.items())
    for k in list(out.keys()):
        if k.endswith("running_mean"):
            p = k[: -len("running_mean")]
            n = out[k].numel()
            out[p + "running_mean"] = torch.randn(n, generator=g) * 0.1
            out[p + "running_var"] = torch.rand(n, generator=g) * 1.0 + 0.5
            lo, hi = (0.3, 0.6) if p.endswith("bn3.") else (0.8, 1.2)
            out[p + "weight"] = torch.rand(n, generator=g) * (hi - lo) + lo
            out[p + "bias"] = torch.randn(n, generator=g) * 0.05
    out["fc.weight"] = torch.randn(out["fc.weight"].shape, generator=g) * fc_std
    out["fc.bias"] = torch.randn(out["fc.bias"].shape, generator=g) * 0.1
    return out


def pooled_features(sd: Dict[str, torch.Tensor], x: torch.Tensor, num_segments: int = 8) -> torch.Tensor:
    """[N, 2048] consensus features (mean over 7x7 and over the segments) the fc acts on, tsm.py:411-419."""
    got = {}
    tsm_forward(sd, x, num_segments=num_segments, tap=lambda n, t: got.__setitem__(n, t))
    f = got["layer4.2.conv3"].mean(dim=(2, 3))
    return f.view(-1, num_segments, f.shape[1]).mean(dim=1)


def fit_head(sd: Dict[str, torch.Tensor], feats: torch.Tensor, targets: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Test-only: replace fc by the least-norm linear head with fc(feats[s]) == targets[s] (feats [S,2048], targets
    [S,C], S <= 2048). Lets a fixture make chosen clips land in chosen states with a known margin — e.g. the two
    poses of a synthetic repetition in states 2k / 2k+1 — without any training."""
    m = feats.mean(dim=0, keepdim=True)
    tm = targets.mean(dim=0, keepdim=True)
    w = (torch.linalg.pinv((feats - m).double()) @ (targets - tm).double()).t().to(torch.float32)   # [C,2048]
    out = OrderedDict((k, v.clone()) for k, v in sd.items())
    out["fc.weight"] = w.contiguous()
    out["fc.bias"] = (tm - m @ w.t()).flatten().contiguous()
    return out


def _conv_key(sd, prefix: str) -> str:
    return prefix + ".net.weight" if prefix + ".net.weight" in sd else prefix + ".weight"


# --------------------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------------------
def temporal_shift(x: torch.Tensor, n_segment: int, fold_div: int) -> torch.Tensor:
    """workoutdetector/models/tsm.py:34-50 (TemporalShift.shift, non-inplace branch)."""
    nt, c, h, w = x.shape
    x5 = x.view(nt // n_segment, n_segment, c, h, w)
    fold = c // fold_div
    left = torch.cat([x5[:, 1:, :fold], torch.zeros_like(x5[:, :1, :fold])], dim=1)            # t <- t+1
    right = torch.cat([torch.zeros_like(x5[:, :1, fold:2 * fold]), x5[:, :-1, fold:2 * fold]], dim=1)  # t <- t-1
    return torch.cat([left, right, x5[:, :, 2 * fold:]], dim=2).reshape(nt, c, h, w)


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, BN_EPS)


def _fold(sd, wkey, bn):
    """BN folded into the conv in fp32: w' = w * g/sqrt(var+eps), b' = beta - mean * g/sqrt(var+eps)."""
    s = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + BN_EPS)
    return sd[wkey] * s.view(-1, 1, 1, 1), sd[bn + ".bias"] - sd[bn + ".running_mean"] * s


def _r(x, emulate):
    return x.to(torch.bfloat16).to(torch.float32) if emulate else x


def tsm_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, num_segments: int = 8, shift_div: int = 8,
                is_shift: bool = True, emulate_bf16: bool = False,
                tap: Optional[Callable[[str, torch.Tensor], None]] = None) -> torch.Tensor:
    """TSM.forward (workoutdetector/models/tsm.py:409-419) over the torchvision ResNet-50 v1.5 trunk it wraps
    (tsm.py:264-283): x [N*T,3,224,224] fp32 -> raw consensus logits [N, num_class].

    emulate_bf16=False: the reference arithmetic (conv, then BatchNorm in eval mode, fp32 throughout).
    emulate_bf16=True : what the bf16 engine computes — BN folded into the weights in fp32, weights and every
    layer output rounded to bf16, fp32 accumulation — so a kernel bug is distinguishable from rounding.
    ``tap(name, tensor)`` receives each op's output (NCHW frames) under the engine's op names."""
    sd = {k: v.to(torch.float32) for k, v in sd.items() if v.is_floating_point()}
    emu = emulate_bf16

    def conv_bn(x, wkey, bn, stride, pad, relu, residual=None):
        if emu:
            w, b = _fold(sd, wkey, bn)
            y = F.conv2d(x, _r(w, True), b, stride=stride, padding=pad)
        else:
            y = _bn(sd, bn, F.conv2d(x, sd[wkey], None, stride=stride, padding=pad))
        if residual is not None:
            y = y + residual
        if relu:
            y = F.relu(y)
        return _r(y, emu)

    def emit(name, t):
        if tap is not None:
            tap(name, t)

    x = _r(x.to(torch.float32), emu)
    x = conv_bn(x, "base_model.conv1.weight", "base_model.bn1", 2, 3, True)
    emit("conv1", x)
    x = F.max_pool2d(x, 3, 2, 1)
    emit("maxpool", x)
    for L in range(4):
        for b in range(BLOCKS[L]):
            p = f"base_model.layer{L + 1}.{b}"
            n = f"layer{L + 1}.{b}"
            stride = 2 if (L > 0 and b == 0) else 1
            xin = temporal_shift(x, num_segments, shift_div) if is_shift else x   # tsm.py:30-32, conv1 only
            y = conv_bn(xin, _conv_key(sd, p + ".conv1"), p + ".bn1", 1, 0, True)
            emit(n + ".conv1", y)
            y = conv_bn(y, p + ".conv2.weight", p + ".bn2", stride, 1, True)
            emit(n + ".conv2", y)
            if b == 0:
                idt = conv_bn(x, p + ".downsample.0.weight", p + ".downsample.1", stride, 0, False)
                emit(n + ".downsample", idt)
            else:
                idt = x                                                            # identity sees the UNshifted x
            x = conv_bn(y, p + ".conv3.weight", p + ".bn3", 1, 0, True, residual=idt)
            emit(n + ".conv3", x)
    f = x.mean(dim=(2, 3))                                                         # AdaptiveAvgPool2d(1), tsm.py:411
    o = F.linear(f, sd["fc.weight"], sd["fc.bias"])                                # per-frame fc, tsm.py:413
    o = o.view(-1, num_segments, o.shape[1]).mean(dim=1)                           # avg consensus, tsm.py:417-419
    return o


def scores_to_states(logits: torch.Tensor, threshold: float = 0.5, softmax: bool = True
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """to_softmax (utils/visualize.py:140-150) + first-max arg-max / threshold (utils/eval.py:159-164)."""
    probs = F.softmax(logits.to(torch.float32), dim=1)
    score = probs if softmax else logits
    top, arg = score.max(dim=1)  # torch.max returns the first maximal index on CPU
    # guard the documented first-max rule explicitly
    arg = (score == top.unsqueeze(1)).to(torch.int64).argmax(dim=1)
    state = torch.where(top >= threshold, arg, torch.full_like(arg, -1))
    return probs, state.to(torch.int32)


# --------------------------------------------------------------------------------------------------
# preprocessing + windowing
# --------------------------------------------------------------------------------------------------
def resize_geometry(H: int, W: int) -> Tuple[int, int, int, int]:
    """torchvision Resize(256) + CenterCrop(224): (rh, rw, top, left). Short side -> 256, long side int(256*l/s);
    crop offsets int(round((size-224)/2.0))."""
    if H <= W:
        rh, rw = 256, int(256 * W / H)
    else:
        rh, rw = int(256 * H / W), 256
    return rh, rw, int(round((rh - 224) / 2.0)), int(round((rw - 224) / 2.0))


def preprocess_u8(frames: torch.Tensor, in_scale: float = 1.0 / 255.0) -> torch.Tensor:
    """build_test_transform(person_crop=False) (workoutdetector/datasets/build.py:131-136) on uint8 HWC frames:
    ConvertImageDtype(float32) [x/255] -> Resize(256) bilinear, align_corners=False, no antialias (the pinned
    torchvision 0.13 tensor path) -> CenterCrop(224) -> Normalize.  frames [n,H,W,3] uint8 -> [n,3,224,224] fp32.
    in_scale=1.0 reproduces the float-promotion quirk of utils/inference_count.py:413-414."""
    n, H, W, _ = frames.shape
    x = frames.permute(0, 3, 1, 2).to(torch.float32) * in_scale
    rh, rw, top, left = resize_geometry(H, W)
    x = F.interpolate(x, size=(rh, rw), mode="bilinear", align_corners=False, antialias=False)
    x = x[:, :, top:top + 224, left:left + 224]
    mean = torch.tensor(MEAN).view(1, 3, 1, 1)
    std = torch.tensor(STD).view(1, 3, 1, 1)
    return (x - mean) / std


def window_indices(total_frames: int, stride: int = 8, span: int = 16, step: int = 2) -> List[List[int]]:
    """Frame indices of every window of inference_dataset (utils/inference_count.py:411-414): windows start at
    0, 8, 16, ...; a window reads frames i, i+2, ..., i+14 that exist; missing ones are -1 (all-zero raw frame)."""
    out = []
    for i in range(0, total_frames, stride):
        idx = [j for j in range(i, min(i + span, total_frames), step)]
        idx += [-1] * (span // step - len(idx))
        out.append(idx)
    return out


def queue_indices(total_frames: int) -> List[List[int]]:
    """count_by_video_model (utils/inference_count.py:313-329): non-overlapping 8-frame queues; a trailing partial
    queue is dropped."""
    return [list(range(i, i + 8)) for i in range(0, total_frames - 7, 8)]
