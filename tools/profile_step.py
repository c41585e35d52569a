"""Minimal program for ncu: one warm-up and one measured forward of the hot path at batch 64 (45 launches each:
preprocess, stem+max-pool, 42 conv kernels, head). Usage under ncu: see profiles/README.md.
WD_NO_SHIFT=1 builds the same network without TemporalShift (is_shift=False): the A/B twin for the DRAM-byte comparison
of the shift-fused conv1 loads (same kernels, same shapes, box t-coordinate 0 instead of +-1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from workoutdetector_b200.models import create_model  # noqa: E402
from workoutdetector_b200.utils.synth import synth_clips_u8  # noqa: E402

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
model = create_model(num_class=12, device="cuda", is_shift=os.environ.get("WD_NO_SHIFT") != "1")
eng = model.engine(clips)
u8 = synth_clips_u8(8, 1).repeat(clips // 8, 1, 1, 1).cuda()
for _ in range(2):
    frames = eng.preprocess_u8(u8)
    logits, probs, state = eng.forward(frames)
torch.cuda.synchronize()
print("ok", float(logits.abs().sum()))
