"""Drop-in for ``workoutdetector.models.tsm`` (reference: workoutdetector/models/tsm.py).

Same constructor / ``create_model`` signature, same ``state_dict`` key layout (so reference checkpoints load
unchanged), same ``forward`` contract — ``[N*T,3,224,224] -> [N, num_class]`` raw consensus logits — but the
forward runs on the sm_100a engine (csrc/) instead of torchvision + cuDNN.  The torch modules below are parameter
containers only: no convolution is ever executed by PyTorch, and there is no CPU path.
"""
import os
import warnings
from collections import OrderedDict
from typing import Optional

import torch
import torchvision
from torch import nn
from torch.nn.init import constant_, normal_

from ..engine import Engine
from . import _sync


class TemporalShift(nn.Module):
    """Parameter container mirroring tsm.py:17-50 (keeps the ``.net`` level in parameter names). The shift itself
    is fused into the A-operand load of the wrapped conv inside the engine."""

    def __init__(self, net: nn.Module, n_segment: int = 3, n_div: int = 8, inplace: bool = False):
        super().__init__()
        self.net = net
        self.n_segment = n_segment
        self.fold_div = n_div
        self.inplace = inplace

    def forward(self, x):  # pragma: no cover - never on the product path
        raise RuntimeError("TemporalShift is fused into the engine's conv1 load; call the TSM module instead")


def make_temporal_shift(net: nn.Module, n_segment: int, n_div: int = 8, place: str = "blockres",
                        temporal_pool: bool = False):
    """tsm.py:104-139 for the supported case (torchvision ResNet, place='blockres', no temporal pool)."""
    if temporal_pool:
        raise NotImplementedError("temporal_pool is hard-wired False in the reference (tsm.py:231)")
    if not isinstance(net, torchvision.models.ResNet):
        raise NotImplementedError(place)
    if "blockres" not in place:
        raise NotImplementedError("only shift_place='blockres' runs on the B200 engine")
    n_round = 2 if len(list(net.layer3.children())) >= 23 else 1
    for j in range(1, 5):
        blocks = list(getattr(net, f"layer{j}").children())
        for i, b in enumerate(blocks):
            if i % n_round == 0:
                blocks[i].conv1 = TemporalShift(b.conv1, n_segment=n_segment, n_div=n_div)
        setattr(net, f"layer{j}", nn.Sequential(*blocks))


def _resnet(base_model: str) -> nn.Module:
    """The reference calls torchvision.models.<base_model>(pretrained=True) (tsm.py:268). There is no network
    here: ImageNet weights are used when they are already in the torch hub cache, else the torchvision init."""
    ctor = getattr(torchvision.models, base_model)
    net = ctor(weights=None)
    try:
        w = torchvision.models.get_model_weights(base_model).IMAGENET1K_V1
        path = os.path.join(torch.hub.get_dir(), "checkpoints", os.path.basename(w.url))
        if os.path.isfile(path):
            net.load_state_dict(torch.load(path, map_location="cpu"))
    except Exception as exc:  # cache lookup is best effort
        warnings.warn(f"ImageNet weights for {base_model} not loaded: {exc}")
    return net


class TSM(nn.Module):
    """TSN with temporal shift module, ResNet-50 / 8 segments / shift_div 8 / 'blockres' on the B200 engine.

    Input ``(batch*num_segments, 3, 224, 224)`` (a 5-D ``(batch, num_segments, 3, 224, 224)`` tensor, which the
    reference's exporters pass but its module rejects, is flattened). Output ``(batch, num_class)`` logits.
    Args as in the reference (tsm.py:189-262).
    """

    def __init__(self, num_class, num_segments=8, base_model="resnet50", consensus_type="avg", before_softmax=True,
                 dropout=0.5, img_feature_dim=256, partial_bn=True, is_shift=True, shift_div=8,
                 shift_place="blockres", fc_lr5=False, non_local=False):
        super().__init__()
        if base_model != "resnet50":
            raise NotImplementedError("the B200 engine implements base_model='resnet50'")
        if consensus_type != "avg":
            raise NotImplementedError("the B200 engine implements consensus_type='avg'")
        if not before_softmax:
            raise NotImplementedError("before_softmax=False: take softmax of the returned logits instead")
        if num_segments != 8:
            raise NotImplementedError("the B200 engine's activation layout is built for num_segments=8")
        self.num_class = num_class
        self.num_segments = num_segments
        self.before_softmax = before_softmax
        self.consensus_type = consensus_type
        self.img_feature_dim = img_feature_dim
        self.temporal_pool = False
        self.is_shift = is_shift
        self.shift_div = shift_div
        self.shift_place = shift_place
        self.fc_lr5 = fc_lr5
        self.non_local = non_local
        self._enable_pbn = partial_bn

        # Same order of RNG draws as the reference constructor, so the same torch.manual_seed gives the same
        # parameters: resnet50 init, fc (tsm.py:246-248), fc again (tsm.py:259-262).
        net = _resnet(base_model)
        if is_shift:
            make_temporal_shift(net, num_segments, n_div=shift_div, place=shift_place)
        self.base_model = net  # registered first so state_dict order matches the reference (tsm.py:268)
        self.input_size = 224
        self.input_mean = [0.485, 0.456, 0.406]
        self.input_std = [0.229, 0.224, 0.225]
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.dropout = nn.Dropout(p=dropout)
        feature_dim = net.fc.in_features
        self.fc = nn.Linear(feature_dim, num_class)
        normal_(self.fc.weight, 0, 0.001)
        constant_(self.fc.bias, 0)
        self.base_model = nn.Sequential(OrderedDict(list(net.named_children())[:-2]))
        self.fc = nn.Linear(feature_dim, num_class)
        normal_(self.fc.weight, 0, 0.001)
        constant_(self.fc.bias, 0)

        self._engine: Optional[Engine] = None
        self._engine_dirty = True
        self._engine_state = None
        self._engine_mode = os.environ.get("WD_B200_MODE", "bf16")
        self._max_clips = 8

    # ---- engine management ---------------------------------------------------------------------
    def set_engine_mode(self, mode: str):
        """'bf16' (product path) or 'fp32' (slow validation path)."""
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
        dev = self.fc.weight.device
        if dev.type != "cuda":
            raise RuntimeError("TSM (B200) runs on a CUDA device only: call .to('cuda') / create_model(device='cuda')")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if self._engine is None or n_clips > self._engine.max_clips or self._engine.device.index != idx:
            self._drop_engine()
            self._max_clips = max(self._max_clips, n_clips)
            self._engine = Engine(self.num_class, max_clips=self._max_clips, mode=self._engine_mode, device=idx,
                                  is_shift=self.is_shift, shift_div=self.shift_div, num_segments=self.num_segments)
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
        """The reference override (tsm.py:285-299) forgets to return self, so ``model.eval()`` yields None there.
        BN is always in eval mode on the engine; this only flips the flag and returns self."""
        if mode:
            warnings.warn("workoutdetector_b200.TSM is inference-only: train(True) has no effect on forward()")
        super().train(mode)
        return self

    def partialBN(self, enable):
        self._enable_pbn = enable

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 5:
            x = x.reshape((-1,) + tuple(x.shape[2:]))
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 224, 224) or x.shape[0] % self.num_segments != 0:
            raise ValueError(f"expected [N*{self.num_segments},3,224,224], got {tuple(x.shape)}")
        n = x.shape[0] // self.num_segments
        eng = self.engine(n)
        x = x.to(eng.device, torch.float32)
        logits, _, _ = eng.forward(eng.pack_nchw(x))
        return logits


def create_model(num_class: int = 2, num_segments: int = 8, base_model: str = "resnet50", checkpoint: str = None,
                 device: str = None, fc_lr5: bool = True, is_shift: bool = True, shift_div: int = 8,
                 shift_place: str = "blockres", consensus_type: str = "avg", img_feature_dim: int = 256,
                 non_local: bool = False, **kwargs) -> nn.Module:
    """Same signature and checkpoint key remapping as the reference (tsm.py:422-476); extra kwargs are swallowed
    as there. ``device`` defaults to cuda when available."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    assert consensus_type in ["avg", "identity"]
    assert shift_place in ["blockres", "block"]
    model = TSM(num_class=num_class, num_segments=num_segments, base_model=base_model,
                consensus_type=consensus_type, img_feature_dim=img_feature_dim, is_shift=is_shift,
                shift_div=shift_div, shift_place=shift_place, fc_lr5=fc_lr5, non_local=non_local)
    if checkpoint is not None:
        ckpt = torch.load(checkpoint, map_location="cpu")
        state_dict = ckpt["state_dict"]
        keys = list(state_dict.keys())
        fc_layer_weight, fc_layer_bias = keys[-2], keys[-1]            # tsm.py:453-454
        if state_dict[fc_layer_weight].shape[0] == num_class:          # tsm.py:455-458
            state_dict["module.fc.weight"] = state_dict[fc_layer_weight]
            state_dict["module.fc.bias"] = state_dict[fc_layer_bias]
        del state_dict[fc_layer_weight]
        del state_dict[fc_layer_bias]
        base_dict = OrderedDict((".".join(k.split(".")[1:]), v) for k, v in state_dict.items())  # tsm.py:462-463
        model.load_state_dict(base_dict, strict=False)
    model.to(device)
    return model
