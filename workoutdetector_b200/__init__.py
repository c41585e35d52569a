"""B200-native (sm_100a) implementation of the WorkoutDetector inference hot path.

Mirrors the reference package layout for the path only: ``models`` (build_model / create_model / TSM),
``utils.inference_count`` (pred_to_count, inference_video, count_by_video_model, inference_dataset, eval_dataset),
``utils.eval`` (main, obo_mae, analyze_count), ``datasets`` (build_test_transform, RepcountHelper).
All device work goes through ``csrc/libwd_b200.so`` (C ABI in include/wd_b200.h); nothing falls back to the CPU.
"""
__version__ = "0.1.0"
