"""Times the non-conv kernels of a step at batch 64: preprocess, counter."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from workoutdetector_b200.engine import count_reps
from workoutdetector_b200.models import create_model
from workoutdetector_b200.utils.synth import synth_clips_u8
B = 64
model = create_model(num_class=12, device="cuda")
eng = model.engine(B)
u8 = synth_clips_u8(8, 1).repeat(B // 8, 1, 1, 1).cuda()
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("preprocess_u8 512 frames: %.1f us" % timeit(lambda: eng.preprocess_u8(u8)))
frames = eng.preprocess_u8(u8)
print("forward: %.1f us" % timeit(lambda: eng.forward(frames)))
st = torch.zeros(1, B, dtype=torch.int32, device="cuda"); lens = torch.full((1,), B, dtype=torch.int32, device="cuda")
print("count_reps: %.1f us" % timeit(lambda: count_reps(st, lens, 8)))
print("torch.empty frames alloc: %.1f us" % timeit(lambda: torch.empty((512,) + eng.frame_shape, dtype=torch.bfloat16, device="cuda")))
