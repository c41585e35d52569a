from .tsm import TSM, create_model  # noqa: F401
from .tdn import TSN  # noqa: F401
from .build import build_model, MODEL_REGISTRY  # noqa: F401
