"""Drop-in for ``workoutdetector.models.build`` (reference: workoutdetector/models/build.py:1-31)."""
import torch


class _Registry(dict):
    """Minimal stand-in for fvcore.common.registry.Registry (declared but unused in the reference, build.py:4)."""

    def __init__(self, name):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        def deco(o):
            self[o.__name__] = o
            return o
        return deco if obj is None else deco(obj)

    def get(self, name):
        if name not in self:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self[name]


MODEL_REGISTRY = _Registry("MODEL")


def build_model(cfg) -> torch.nn.Module:
    """cfg.model.model_type (case-insensitive) selects the constructor; ``**cfg.model`` is forwarded exactly as the
    reference does (build.py:23-31). Both 'tsm' and 'tdn' run on the B200 engine."""
    model_type = cfg.model.model_type.lower()
    if model_type == "tsm":
        from .tsm import create_model as create_model_tsm
        return create_model_tsm(**cfg.model)
    if model_type == "tdn":
        from .tdn import create_model as create_model_tdn
        return create_model_tdn(**cfg.model)
    raise KeyError(f"Model '{cfg.model.model_type}' is not supported.")
