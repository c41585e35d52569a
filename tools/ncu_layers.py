"""Tiny program for `ncu --set full`: a few single-layer launches (names from gpu_bench_layers.LAYERS) at batch 64.
Each layer is launched 1 (check) + 2 (warm-up) + 1 (timed) times; pick the launch with ncu -s/-c.
Usage: python tools/ncu_layers.py <version> <layer-name-substring>[:mode] ..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gpu_bench_layers import LAYERS  # noqa: E402
from workoutdetector_b200.engine import bench_conv  # noqa: E402

ver = int(sys.argv[1])
clips = int(os.environ.get("WD_CLIPS", "64"))
for want in sys.argv[2:]:
    want, _, md = want.partition(":")
    for name, H, Cin, Cout, k, stride, fold, res, mode, tn in LAYERS:
        if name == want:
            ms = bench_conv(clips, H, Cin, Cout, k, stride, fold, bool(res), md or mode, tn, ver, 1)
            print(f"{name} {md or mode} v{ver + 1}: {ms * 1e3:.1f} us", flush=True)
