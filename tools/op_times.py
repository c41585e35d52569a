"""Per-op device times of one forward at batch B (CUDA events around every launch, wd_forward_timed), with the
algorithmic bytes (input + residual + output once, bf16) and FLOPs of each op. Usage: python tools/op_times.py [B] [iters]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from workoutdetector_b200.models import create_model  # noqa: E402
from workoutdetector_b200.utils.synth import synth_clips_u8  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
opts = dict(kv.split("=") for kv in os.environ.get("WD_OPTS", "").split(",") if kv)
torch.manual_seed(0)
model = create_model(num_class=12, device="cuda")
eng = model.engine(B)
if opts:
    for k, v in opts.items():
        eng.set_option(k, int(v))
    eng.load_state_dict(model.state_dict())
u8 = synth_clips_u8(8, 1).repeat((B + 7) // 8, 1, 1, 1)[: B * 8].cuda()
frames = eng.preprocess_u8(u8)
ops = eng.ops()
acc = [0.0] * len(ops)
for i in range(iters + 2):
    *_, ms = eng.forward(frames, timed=True)
    if i >= 2:
        acc = [a + m for a, m in zip(acc, ms)]
ms = [a / iters for a in acc]
tot = sum(ms)
rows = []
prev_c, prev_hw = 4, 224 * 240
for o, m in zip(ops, ms):
    hw = o["hout"] * o["wout"]
    if o["kind"] == "head":
        byts = B * 8 * prev_hw * prev_c * 2
    elif o["kind"] == "stem_pool":
        byts = B * 8 * (224 * 240 * 4 + hw * o["cout"]) * 2
    else:
        hin = hw * o["stride"] ** 2 if o["kind"] == "conv" else hw * 4
        cin = o["cin"] if o["kind"] == "conv" else o["cout"]
        res = 1 if o["name"].endswith("conv3") else 0
        byts = B * 8 * (hin * cin / (o["stride"] ** 2 if o["k"] == 1 else 1) + hw * o["cout"] * (1 + res)) * 2
    fl = 2.0 * o["macs_per_clip"] * B
    rows.append(dict(name=o["name"], kind=o["kind"], a_mode=o["a_mode"], tile_n=o["tile_n"], ms=m,
                     tflops=fl / m / 1e9, gbs=byts / m / 1e6))
    print(f"{o['name']:22s} {o['kind']:9s} {o['a_mode']:6s} n{o['tile_n']:<3d} {m * 1e3:8.1f} us  {fl / m / 1e9:7.1f} TF/s "
          f"{byts / m / 1e6:7.0f} GB/s  {100 * m / tot:4.1f}%")
    prev_c, prev_hw = o["cout"], hw
print(f"sum of ops {tot:.3f} ms  -> {B / tot * 1e3:.0f} clips/s")
out = os.environ.get("WD_OUT")
if out:
    with open(out, "w") as f:
        json.dump(dict(batch=B, ms=tot, ops=rows), f, indent=1)
# whole-forward loops (no per-op events): forward alone, preprocess + forward
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def loop(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0.record()
    for _ in range(n):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n


print(f"forward alone          {loop(lambda: eng.forward(frames)):.3f} ms")
print(f"preprocess + forward   {loop(lambda: eng.forward(eng.preprocess_u8(u8))):.3f} ms")
