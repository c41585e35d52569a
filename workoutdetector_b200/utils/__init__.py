from .inference_count import (COLORS, count_by_video_model, eval_dataset, inference_dataset,  # noqa: F401
                              inference_video, pred_to_count, pred_to_count_batch)
from .eval import analyze_count, obo_mae, to_softmax  # noqa: F401
