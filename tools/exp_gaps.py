"""Where do the inter-kernel gaps come from?  Whole-forward loops at batch B with PDL on/off, short and long loops,
with the SM clock sampled through NVML meanwhile.  Usage: python tools/exp_gaps.py [B]"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from workoutdetector_b200.models import create_model  # noqa: E402
from workoutdetector_b200.utils.synth import synth_clips_u8  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
model = create_model(num_class=12, device="cuda")
eng = model.engine(B)
u8 = synth_clips_u8(8, 1).repeat((B + 7) // 8, 1, 1, 1)[: B * 8].cuda()
frames = eng.preprocess_u8(u8)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

clk = []
stop = False


def sampler():
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(0)
    while not stop:
        clk.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(h) / 1e3))
        time.sleep(0.005)


th = threading.Thread(target=sampler, daemon=True)
th.start()


def loop(n, timed=False):
    for _ in range(3):
        eng.forward(frames)
    torch.cuda.synchronize()
    a = time.perf_counter()
    t0.record()
    for _ in range(n):
        eng.forward(frames, timed=True) if timed else eng.forward(frames)
    t1.record()
    torch.cuda.synchronize()
    b = time.perf_counter()
    s = [c for (t, c, p) in clk if a <= t <= b]
    pw = [p for (t, c, p) in clk if a <= t <= b]
    return t0.elapsed_time(t1) / n, (min(s), sorted(s)[len(s) // 2], max(s)) if s else None, max(pw) if pw else None


for pdl in (1, 0, 1):
    eng.set_option("pdl", pdl)
    for n in (5, 20, 100):
        ms, c, pw = loop(n)
        print(f"pdl={pdl} n={n:3d}: {ms:.3f} ms per forward, SM MHz min/med/max {c}, max power {pw} W", flush=True)
ms, c, pw = loop(5, timed=True)
print(f"per-op events (timed=True) n=5: {ms:.3f} ms per forward, SM MHz {c}, max power {pw} W")
stop = True
