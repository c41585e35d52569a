"""Per-op device times of one TDN-R50 forward at batch B (BASELINE configs[4]: batch 128, 11 classes), CUDA events
around every op (wd_forward_timed).  Usage: python tools/op_times_tdn.py [B] [iters]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from workoutdetector_b200.models.tdn import create_model  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.manual_seed(0)
model = create_model(num_class=11).to("cuda")
eng = model.engine(B)
g = torch.Generator(device="cuda").manual_seed(1)
chunk = 16
x = torch.randn(chunk, 8, 5, 3, 224, 224, device="cuda", generator=g)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
clips = eng.pack_tdn(x)
reps = (B + chunk - 1) // chunk
clips_all = torch.empty(B * eng.clip_bytes // 2, dtype=torch.bfloat16, device="cuda")
# assemble a B-clip buffer: centre frames of all clips first, then all difference tensors
fr = 8 * eng.frame_shape[0] * eng.frame_shape[1] * eng.frame_shape[2]
df = eng.clip_bytes // 2 - fr
for r in range(reps):
    n = min(chunk, B - r * chunk)
    clips_all[r * chunk * fr:(r * chunk + n) * fr] = clips[:n * fr]
    clips_all[B * fr + r * chunk * df:B * fr + (r * chunk + n) * df] = clips[chunk * fr:chunk * fr + n * df]
t0.record()
eng.pack_tdn(x)
t1.record()
torch.cuda.synchronize()
print(f"pack_tdn of {chunk} clips: {t0.elapsed_time(t1):.3f} ms")
ops = eng.ops()
acc = [0.0] * len(ops)
for i in range(iters + 2):
    *_, ms = eng.forward(clips_all, timed=True)
    if i >= 2:
        acc = [a + m for a, m in zip(acc, ms)]
ms = [a / iters for a in acc]
tot = sum(ms)
by_kind = {}
for o, m in zip(ops, ms):
    fl = 2.0 * o["macs_per_clip"] * B
    by_kind[o["kind"]] = by_kind.get(o["kind"], 0.0) + m
    print(f"{o['name']:22s} {o['kind']:9s} {o['a_mode']:6s} n{o['tile_n']:<3d} {m * 1e3:8.1f} us  {fl / m / 1e9:7.1f} TF/s "
          f"{100 * m / tot:4.1f}%")
gflop = sum(2.0 * o["macs_per_clip"] for o in ops) / 1e9
print("by kind:", {k: round(v, 3) for k, v in by_kind.items()})
print(f"sum of ops {tot:.3f} ms -> {B / tot * 1e3:.0f} clips/s; {gflop:.2f} GFLOP/clip -> {B * gflop / tot:.0f} TFLOP/s")
# end to end: wd_forward without per-op events
for _ in range(2):
    eng.forward(clips_all)
t0.record()
for _ in range(iters):
    eng.forward(clips_all)
t1.record()
torch.cuda.synchronize()
e2e = t0.elapsed_time(t1) / iters
print(f"wd_forward: {e2e:.3f} ms per {B} clips -> {B / e2e * 1e3:.0f} clips/s")
out = os.environ.get("WD_OUT")
if out:
    with open(out, "w") as f:
        json.dump(dict(batch=B, ms=tot, forward_ms=e2e, by_kind=by_kind,
                       ops=[dict(name=o["name"], kind=o["kind"], ms=m) for o, m in zip(ops, ms)]), f, indent=1)
