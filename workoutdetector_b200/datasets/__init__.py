from .repcount_dataset import RepcountHelper, RepcountItem, RepcountItemWithPred, eval_count  # noqa: F401
from .build import build_test_transform, MEAN_STD, INPUT_SIZE  # noqa: F401
