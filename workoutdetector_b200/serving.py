"""Serving boundary: the engine behind the onnxruntime.InferenceSession call surface the
reference's callers use — ``inference_video`` (utils/inference_count.py:265-276), ``app/inference.py:54-78`` and
``demo.py:82-109`` all do ``name = sess.get_inputs()[0].name; out = sess.run(None, {name: x})[0]`` with
``x`` = float32 ndarray [N, 8, 3, 224, 224] (normalised) and expect an ndarray [N, num_class].

``EngineSession(model)`` wraps a workoutdetector_b200 TSM / TSN(TDN) module; ``softmax=True`` gives the mmaction2-style
probability output of the 11-class action-recognition demo (configs/tsm_action_recogition_sthv2.py: average_clips='prob').
"""
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


class EngineSession:
    def __init__(self, model: torch.nn.Module, softmax: bool = False, input_name: str = "input",
                 output_name: str = "output"):
        if not hasattr(model, "engine"):
            raise TypeError("EngineSession wraps a workoutdetector_b200 model (TSM or TSN/TDN)")
        self.model = model
        self.softmax = softmax
        self._in = input_name
        self._out = output_name
        self._tdn = model.__class__.__name__ == "TSN"

    def get_inputs(self) -> List[SimpleNamespace]:
        shape = ["N", 8, 5, 3, 224, 224] if self._tdn else ["N", 8, 3, 224, 224]
        return [SimpleNamespace(name=self._in, shape=shape, type="tensor(float)")]

    def get_outputs(self) -> List[SimpleNamespace]:
        return [SimpleNamespace(name=self._out, shape=["N", self.model.num_class], type="tensor(float)")]

    def run(self, output_names: Optional[Sequence[str]], input_feed: Dict[str, np.ndarray]) -> List[np.ndarray]:
        if self._in not in input_feed:
            raise KeyError(f"input '{self._in}' missing from the feed (got {list(input_feed)})")
        x = torch.from_numpy(np.ascontiguousarray(input_feed[self._in], dtype=np.float32))
        dev = next(self.model.parameters()).device
        with torch.no_grad():
            y = self.model(x.to(dev))
            if self.softmax:
                y = torch.softmax(y, dim=1)
        return [y.cpu().numpy()]
