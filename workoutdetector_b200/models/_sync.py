"""Keeps an engine's packed weights in step with the torch parameter container that owns them.

The engine holds BN-folded, bf16-packed copies of the module's parameters.  ``load_state_dict`` / ``.to()`` flag the
module explicitly; every other way of changing a weight — ``model.fc.weight.data.copy_(...)``, an optimizer step,
``model.base_model.load_state_dict(...)``, replacing ``model.fc`` — is caught here, in two tiers:

  * a fingerprint of (identity, data_ptr, _version) over all parameters and buffers: a few microseconds per call, sees
    replaced tensors, moved storage and every in-place op autograd knows about;
  * a content checksum (one multi-tensor ``_foreach_norm`` launch + one scalar read-back): sees writes through ``.data``,
    which carry their own version counter.  On by default on the nn.Module path (a 1-clip forward costs ~1 ms, the check
    ~0.1 ms); ``WD_B200_WEIGHT_CHECK=0`` turns it off for callers that promise to call ``refresh_engine()`` themselves.
The hot paths that drive the engine object directly (bench.py, utils.inference_count.score_windows after its first
call) do not pay for either more than once per ``model.engine()`` call.
"""
import itertools
import os
from typing import Tuple

import torch
from torch import nn

_CHECK = os.environ.get("WD_B200_WEIGHT_CHECK", "1") != "0"


def fingerprint(module: nn.Module) -> Tuple[int, int]:
    acc, n = 0, 0
    for t in itertools.chain(module.parameters(), module.buffers()):
        acc = (acc * 1000003 + (id(t) ^ (t.data_ptr() * 31) ^ (t._version * 131071))) & 0xFFFFFFFFFFFFFFF
        n += 1
    return acc, n


def checksum(module: nn.Module) -> float:
    """Sum of the L2 norms of every floating-point parameter / buffer, accumulated in float64 on the tensors' device."""
    if not _CHECK:
        return 0.0
    ts = [t.detach() for t in itertools.chain(module.parameters(), module.buffers()) if t.is_floating_point() and t.numel()]
    if not ts:
        return 0.0
    norms = torch._foreach_norm(ts)
    return float(torch.stack([x.double() for x in norms]).sum())


def state_of(module: nn.Module):
    return fingerprint(module), checksum(module)
