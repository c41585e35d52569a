"""Pipeline timeline of CTA 0 of one v4 conv launch (WD_TRACE hook): per-role clock() marks -> wait / work cycles.
Usage: python tools/trace_conv.py "<layer name>[:mode]" [first_marks]"""
import os
import struct
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = "/tmp/wd_trace.bin"
os.environ["WD_TRACE"] = path
from tools.gpu_bench_layers import LAYERS  # noqa: E402
from workoutdetector_b200.engine import bench_conv  # noqa: E402

want, _, md = sys.argv[1].partition(":")
nshow = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for name, H, Cin, Cout, k, stride, fold, res, mode, tn in LAYERS:
    if name == want:
        ms = bench_conv(64, H, Cin, Cout, k, stride, fold, bool(res), md or mode, tn, 3, 1)
        print(f"{name} {md or mode}: {ms * 1e3:.1f} us (the trace is of the first, cold launch)")
raw = open(path, "rb").read()
v = struct.unpack(f"{len(raw) // 4}I", raw)
roles = ["mma", "w_producer", "a_producer", "epilogue(w0)"]
for r, nm in enumerate(roles):
    t = [x for x in v[r * 2048:(r + 1) * 2048] if x]
    if not t:
        continue
    d = [(b - a) & 0xFFFFFFFF for a, b in zip(t, t[1:])]
    print(f"{nm}: {len(t)} marks, span {(t[-1] - t[0]) & 0xFFFFFFFF} cycles; deltas:")
    print("   ", " ".join(str(x) for x in d[:nshow]))
    mid = d[len(d) // 2: len(d) // 2 + nshow]
    print("    mid:", " ".join(str(x) for x in mid))
# absolute timeline of a steady-state window (all roles share the SM clock)
t0 = min(x for x in v if x)
def window(r, per, kb0, n):
    t = [x for x in v[r * 2048:(r + 1) * 2048] if x]
    return [(t[i] - t0) & 0xFFFFFFFF for i in range(kb0 * per, min(len(t), (kb0 + n) * per))]
kb0 = int(os.environ.get("WD_KB0", "40"))
print("absolute cycles, k-blocks", kb0, "..", kb0 + 5)
print("  mma  (before wait_a, after wait_a, after wait_b):", window(0, 3, kb0, 6))
print("  w    (before wait_empty, after):", window(1, 2, kb0, 6))
print("  a    (before wait_empty, after):", window(2, 2, kb0, 6))
print("  epi  (marks per chunk: e0 start, e1 res landed, e2 acc loaded, e3 out slab free, e4 math+STS, e5 fence; +2 per tile around tmem_full):", window(3, 1, 60, 60))
