"""Two TDN-R50 forwards at batch B (one warm-up, one to profile) for ncu launch lists.  Usage under ncu:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tdn_launches.csv \
    python tools/profile_step_tdn.py 128"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from workoutdetector_b200.models.tdn import create_model  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.manual_seed(0)
model = create_model(num_class=11).to("cuda")
eng = model.engine(B)
x = torch.randn(8, 8, 5, 3, 224, 224, device="cuda")
small = eng.pack_tdn(x)
fr = 8 * eng.frame_shape[0] * eng.frame_shape[1] * eng.frame_shape[2]
df = eng.clip_bytes // 2 - fr
clips = torch.empty(B * eng.clip_bytes // 2, dtype=torch.bfloat16, device="cuda")
for r in range((B + 7) // 8):
    n = min(8, B - r * 8)
    clips[r * 8 * fr:(r * 8 + n) * fr] = small[:n * fr]
    clips[B * fr + r * 8 * df:B * fr + (r * 8 + n) * df] = small[8 * fr:8 * fr + n * df]
for _ in range(2):
    eng.forward(clips)
    torch.cuda.synchronize()
print("done")
