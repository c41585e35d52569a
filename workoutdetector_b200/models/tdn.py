"""Drop-in for ``workoutdetector.models.tdn`` + ``workoutdetector.models.tsn`` (reference: models/tdn.py, models/tsn.py).

Same ``create_model`` signature, same ``state_dict`` key names, shapes and order as the reference's ``TSN(TDN_Net)``
module (so its checkpoints load unchanged, and ``load_state_dict(strict=True)`` of a reference state_dict succeeds),
same ``forward`` contract — ``[B, 8, 5, 3, 224, 224]`` or ``[B*40, 3, 224, 224]`` normalised floats -> ``[B, num_class]``
consensus logits — but the forward runs on the sm_100a engine (csrc/).  The torch modules below are parameter
containers only: PyTorch never executes a convolution here and there is no CPU path.

Initial values follow the reference's distributions (FBResNet: N(0, sqrt(2/n)) convolutions, identity BatchNorm,
tdn.py:541-547; conv1_5 = channel-mean of the RGB stem kernel, tdn.py:106-115; ShiftModule 'shift' init,
tdn.py:352-358; new_fc N(0, 0.01), tsn.py:156-158) but not its exact RNG draw order: weights come from checkpoints.
"""
import math
import os
import warnings
from typing import Optional

import torch
from torch import nn
from torch.nn.init import constant_, normal_

from ..engine import Engine
from . import _sync


class mSEModule(nn.Module):
    """Parameters of tdn.py:188-249; the arithmetic is csrc/wd_tdn_kernels.cuh (mse_* kernels)."""

    def __init__(self, channel, n_segment=8, index=1):
        super().__init__()
        self.channel = channel
        self.reduction = 16
        self.n_segment = n_segment
        r = channel // self.reduction
        self.conv1 = nn.Conv2d(channel, r, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm2d(r)
        self.conv2 = nn.Conv2d(r, r, kernel_size=3, padding=1, groups=r, bias=False)
        self.conv3 = nn.Conv2d(r, channel, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(channel)
        self.conv3_smallscale2 = nn.Conv2d(r, r, padding=1, kernel_size=3, bias=False)
        self.bn3_smallscale2 = nn.BatchNorm2d(r)
        self.conv3_smallscale4 = nn.Conv2d(r, r, padding=1, kernel_size=3, bias=False)
        self.bn3_smallscale4 = nn.BatchNorm2d(r)


class ShiftModule(nn.Module):
    """Parameters of tdn.py:339-364 (depthwise temporal Conv1d, 'shift' init)."""

    def __init__(self, input_channels, n_segment=8, n_div=8, mode="shift"):
        super().__init__()
        self.input_channels = input_channels
        self.n_segment = n_segment
        self.fold_div = n_div
        self.fold = input_channels // n_div
        self.conv = nn.Conv1d(input_channels, input_channels, kernel_size=3, padding=1, groups=input_channels,
                              bias=False)
        if mode == "shift":
            self.conv.weight.data.zero_()
            self.conv.weight.data[:self.fold, 0, 2] = 1
            self.conv.weight.data[self.fold:2 * self.fold, 0, 0] = 1
            if 2 * self.fold < input_channels:
                self.conv.weight.data[2 * self.fold:, 0, 1] = 1
        elif mode == "fixed":
            self.conv.weight.data.zero_()
            self.conv.weight.data[:, 0, 1] = 1


class Bottleneck(nn.Module):
    """tdn.py:410-457 (plain) / 460-520 (``shift=True``: with mSEModule + ShiftModule); convolutions have bias=True."""
    expansion = 4

    def __init__(self, num_segments, inplanes, planes, stride=1, downsample=None, shift=False):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=True)
        self.bn1 = nn.BatchNorm2d(planes)
        if shift:
            self.num_segments = num_segments
            self.mse = mSEModule(planes, n_segment=num_segments, index=1)
            self.shift = ShiftModule(planes, n_segment=num_segments, n_div=8, mode="shift")
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=True)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=True)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


def _make_layer(num_segments, inplanes, planes, blocks, stride, shift):
    """tdn.py:549-567."""
    downsample = None
    if stride != 1 or inplanes != planes * 4:
        downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 4, kernel_size=1, stride=stride, bias=True),
                                   nn.BatchNorm2d(planes * 4))
    layers = [Bottleneck(num_segments, inplanes, planes, stride, downsample, shift)]
    for _ in range(1, blocks):
        layers.append(Bottleneck(num_segments, planes * 4, planes, shift=shift))
    return nn.Sequential(*layers)


class TDN_Net(nn.Module):
    """Parameter container with the attribute names / registration order of tdn.py:92-137 (two FBResNet-50 trunks:
    the RGB one gives conv1, bn1, layer1..4_bak; the second gives conv1_temp, conv1_5 and resnext_layer1)."""

    def __init__(self, num_segments=8, alpha=0.5, beta=0.5):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=True)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.conv1_temp = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=True)
        self.conv1_5 = nn.Sequential(nn.Conv2d(12, 64, kernel_size=7, stride=2, padding=3, bias=False),
                                     nn.BatchNorm2d(64), nn.ReLU(inplace=True))
        self.maxpool_diff = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.resnext_layer1 = _make_layer(num_segments, 64, 64, 3, 1, False)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1_bak = _make_layer(num_segments, 64, 64, 3, 1, False)
        self.layer2_bak = _make_layer(num_segments, 256, 128, 4, 2, True)
        self.layer3_bak = _make_layer(num_segments, 512, 256, 6, 2, True)
        self.layer4_bak = _make_layer(num_segments, 1024, 512, 3, 2, True)
        self.avgpool = nn.AdaptiveAvgPool2d(1)   # tsn.py:169
        self.avg_diff = nn.AvgPool2d(kernel_size=2, stride=2)
        self.fc = nn.Dropout(p=0.5)              # replaced by TSN._prepare_tsn (tsn.py:146-148)
        self.alpha = alpha
        self.beta = beta
        for m in self.modules():                 # tdn.py:541-547
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        self.conv1_5[0].weight.data = self.conv1_temp.weight.data.mean(dim=1, keepdim=True).expand(
            64, 12, 7, 7).contiguous()           # tdn.py:106-115


class TSN(nn.Module):
    """``workoutdetector.models.TSN`` ("Only for TDN", tsn.py:99-135) on the B200 engine: ResNet-50, 8 segments,
    5 frames per segment, average consensus."""

    def __init__(self, num_class: int, num_segments: int = 8, backbone_fn=None, num_frames: int = 5,
                 base_model: str = "resnet50", consensus_type: str = "avg", dropout: float = 0.8,
                 init_std: float = 0.01, partial_bn=True, fc_lr5=False):
        super().__init__()
        assert "resnet" in base_model, ValueError(f"Unknown base model: {base_model}")
        if "50" not in base_model:
            raise NotImplementedError("the B200 engine implements base_model='resnet50'")
        if num_segments != 8:
            raise NotImplementedError("the B200 engine's activation layout is built for num_segments=8")
        if num_frames != 5:
            raise NotImplementedError("TDN takes 5 frames per segment (tdn.py:31)")
        if consensus_type != "avg":
            raise NotImplementedError("the B200 engine implements consensus_type='avg'")
        if dropout == 0:
            raise NotImplementedError("dropout=0 moves the classifier into base_model.fc (tsn.py:141-144)")
        self.num_class = num_class
        self.num_segments = num_segments
        self.reshape = True
        self.dropout = dropout
        self.consensus_type = consensus_type
        self.fc_lr5 = fc_lr5
        self.init_std = init_std
        self.base_model = TDN_Net(num_segments, alpha=0.5, beta=0.5)   # tdn_net(): 0.5 / 0.5 for 8 segments
        self.base_model.fc = nn.Dropout(p=dropout)
        self.input_size = 224
        self.input_mean = [0.485, 0.456, 0.406]
        self.input_std = [0.229, 0.224, 0.225]
        self.new_fc = nn.Linear(2048, num_class)
        normal_(self.new_fc.weight, 0, init_std)
        constant_(self.new_fc.bias, 0)
        self.modality = "RGB"
        self.new_length = num_frames
        self.before_softmax = True
        self._enable_pbn = partial_bn

        self._engine: Optional[Engine] = None
        self._engine_dirty = True
        self._engine_state = None
        self._engine_mode = os.environ.get("WD_B200_MODE", "bf16")
        self._max_clips = 4

    # ---- engine management (as models/tsm.py) -------------------------------------------------------
    def set_engine_mode(self, mode: str):
        assert mode in ("bf16", "fp32")
        if mode != self._engine_mode:
            self._engine_mode = mode
            self._drop_engine()

    def _drop_engine(self):
        if self._engine is not None:
            self._engine.close()
        self._engine = None
        self._engine_dirty = True

    def engine(self, n_clips: int = 1) -> Engine:
        dev = self.new_fc.weight.device
        if dev.type != "cuda":
            raise RuntimeError("TSN/TDN (B200) runs on a CUDA device only: call .to('cuda')")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if self._engine is None or n_clips > self._engine.max_clips or self._engine.device.index != idx:
            self._drop_engine()
            self._max_clips = max(self._max_clips, n_clips)
            self._engine = Engine(self.num_class, max_clips=self._max_clips, mode=self._engine_mode, device=idx,
                                  arch="tdn")
        # weights changed behind the engine's back (in-place edits, optimizer steps, sub-module load_state_dict, a
        # replaced fc)?  models/_sync.py: version / pointer fingerprint + content checksum
        state = _sync.state_of(self)
        if self._engine_dirty or state != self._engine_state:
            self._engine.load_state_dict(self.state_dict())
            self._engine_dirty = False
            self._engine_state = state
        return self._engine

    def refresh_engine(self) -> None:
        """Re-upload the module's current parameters on the next forward, unconditionally."""
        self._engine_dirty = True

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._engine_dirty = True
        return out

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._engine_dirty = True
        return out

    def train(self, mode=True):
        if mode:
            warnings.warn("workoutdetector_b200.TSN is inference-only: train(True) has no effect on forward()")
        super().train(mode)
        return self

    def partialBN(self, enable):
        self._enable_pbn = enable

    def forward(self, input: torch.Tensor, reshape: bool = True) -> torch.Tensor:
        """tsn.py:335-351: any input that reshapes to [-1, 15, 224, 224]; one clip = 8 segments x 5 frames."""
        if tuple(input.shape[-2:]) != (224, 224) or input.numel() % (120 * 224 * 224) != 0:
            raise ValueError(f"expected [B,8,5,3,224,224] or [B*40,3,224,224], got {tuple(input.shape)}")
        n = input.numel() // (120 * 224 * 224)
        eng = self.engine(n)
        logits, _, _ = eng.forward(eng.pack_tdn(input.to(eng.device, torch.float32)))
        return logits


def create_model(num_class: int, num_segments: int = 8, base_model: str = "resnet50", num_frames: int = 5,
                 checkpoint: str = None, consensus_type="avg", dropout: float = 0.5, partial_bn: bool = False,
                 fc_lr5: bool = False, **kwargs) -> nn.Module:
    """Same signature and checkpoint handling as the reference (tdn.py:20-73): ``.net`` infixes are added / removed to
    match, fc weights of a different class count are dropped.  The model is returned on the CPU like the reference's;
    move it with ``.to('cuda')`` before calling it."""
    model = TSN(num_class=num_class, num_segments=num_segments, num_frames=num_frames, base_model=base_model,
                consensus_type=consensus_type, dropout=dropout, partial_bn=partial_bn, fc_lr5=fc_lr5)
    if not checkpoint:
        return model
    print("=> fine-tuning from '{}'".format(checkpoint))
    sd = torch.load(checkpoint, map_location="cpu")["state_dict"]
    fc_layer_weight = list(sd.keys())[-2]
    model_dict = model.state_dict()
    replace = []
    for k in sd:
        if k not in model_dict and k.replace(".net", "") in model_dict:
            replace.append((k, k.replace(".net", "")))
    for k in model_dict:
        if k not in sd and k.replace(".net", "") in sd:
            replace.append((k.replace(".net", ""), k))
    for k, k_new in replace:
        sd[k_new] = sd.pop(k)
    if sd[fc_layer_weight].shape != model_dict["new_fc.weight"].shape:
        print("=> New dataset, do not load fc weights")
        sd = {k: v for k, v in sd.items() if "fc" not in k}
    model_dict.update(sd)
    model.load_state_dict(model_dict)
    return model
