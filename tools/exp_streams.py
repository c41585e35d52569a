"""Experiment: split one batch of B clips over S sub-engines on S streams (tail filling: the persistent one-CTA-per-SM
convolution kernels of one sub-batch fill the wave-quantisation tails and launch gaps of the other).
Usage: python tools/exp_streams.py [B]   (prints ms per B-clip step and clips/s for several splits)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from workoutdetector_b200.engine import Engine  # noqa: E402
from workoutdetector_b200.models import create_model  # noqa: E402
from workoutdetector_b200.utils.synth import synth_clips_u8  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
model = create_model(num_class=12, device="cuda")
sd = model.state_dict()
u8 = synth_clips_u8(8, 1).repeat((B + 7) // 8, 1, 1, 1)[: B * 8].cuda()


def run(split, pdl, steps=20):
    n = B // split
    engs = []
    for _ in range(split):
        e = Engine(12, max_clips=n)
        e.set_option("pdl", pdl)
        e.load_state_dict(sd)
        engs.append(e)
    frames = [engs[i].preprocess_u8(u8[i * n * 8:(i + 1) * n * 8]) for i in range(split)]
    streams = [torch.cuda.Stream() for _ in range(split)]
    main = torch.cuda.current_stream()
    outs = [None] * split

    def step():
        ev = torch.cuda.Event()
        ev.record(main)
        for i in range(split):
            streams[i].wait_event(ev)
            with torch.cuda.stream(streams[i]):
                outs[i] = engs[i].forward(frames[i])
            done = torch.cuda.Event()
            done.record(streams[i])
            main.wait_event(done)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    states = torch.cat([o[2] for o in outs]).cpu()
    for e in engs:
        e.close()
    return ms, states


ref = None
for split, pdl in [(1, 1), (1, 0), (2, 1), (2, 0), (4, 1), (4, 0), (8, 0)]:
    if B % split:
        continue
    ms, st = run(split, pdl)
    if ref is None:
        ref = st
    same = bool((st == ref).all())
    print(f"B={B} split={split} pdl={pdl}: {ms:.3f} ms/step  {B / ms * 1e3:8.0f} clips/s  states_equal={same}", flush=True)
