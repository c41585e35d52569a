"""Single-layer timing on the B200 box: each TSM-R50 conv shape at several batch sizes, so that the effect of L2
residency (small batches fit the 126 MB L2) and of tile shape can be read directly. Prints ms, TFLOP/s and the
effective GB/s over the algorithmic bytes (input + residual + output once)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from workoutdetector_b200.engine import bench_conv  # noqa: E402

LAYERS = [
    # name, H, Cin, Cout, k, stride, fold, residual, mode, tile_n
    ("l1.conv1 256->64", 56, 256, 64, 1, 1, 32, 0, "gather", 64),
    ("l1.conv2 3x3 64", 56, 64, 64, 3, 1, 0, 0, "gather", 64),
    ("l1.conv3 64->256+res", 56, 64, 256, 1, 1, 0, 1, "tma", 256),
    ("l1.conv3 64->256+res n128", 56, 64, 256, 1, 1, 0, 1, "tma", 128),
    ("l2.conv1 512->128", 28, 512, 128, 1, 1, 64, 0, "tma", 128),
    ("l2.conv2 3x3 128", 28, 128, 128, 3, 1, 0, 0, "gather", 128),
    ("l2.conv3 128->512+res", 28, 128, 512, 1, 1, 0, 1, "tma", 256),
    ("l3.conv1 1024->256", 14, 1024, 256, 1, 1, 128, 0, "tma", 256),
    ("l3.conv2 3x3 256", 14, 256, 256, 3, 1, 0, 0, "gather", 256),
    ("l3.conv2 3x3 256 n128", 14, 256, 256, 3, 1, 0, 0, "gather", 128),
    ("l3.conv3 256->1024+res", 14, 256, 1024, 1, 1, 0, 1, "tma", 256),
    ("l4.conv2 3x3 512", 7, 512, 512, 3, 1, 0, 0, "gather", 256),
    ("l4.conv3 512->2048+res", 7, 512, 2048, 1, 1, 0, 1, "tma", 256),
]

if __name__ == "__main__":
    clip_list = [int(c) for c in os.environ.get("WD_CLIPS", "64,16").split(",")]
    versions = [int(v) for v in os.environ.get("WD_VERSIONS", "1,2").split(",")]
    for name, H, Cin, Cout, k, stride, fold, res, mode, tn in LAYERS:
        for clips in clip_list:
            for ver in versions:
                modes = [mode]
                if ver >= 2 and k == 3 and stride == 1 and H % 14 == 0:
                    modes.append("strip")
                for md in modes:
                    try:
                        ms = bench_conv(clips, H, Cin, Cout, k, stride, fold, bool(res), md, tn, ver, 20)
                    except Exception as ex:  # keep going: this is a measurement sweep
                        print(f"{name:28s} clips {clips:3d} v{ver + 1} {md}: FAILED {ex}", flush=True)
                        continue
                    Ho = H // stride
                    flops = 2.0 * clips * 8 * Ho * Ho * Cout * Cin * k * k
                    byts = 2.0 * clips * 8 * (H * H * Cin + Ho * Ho * Cout * (2 if res else 1))
                    print(f"{name:28s} clips {clips:3d} v{ver + 1} {md:6s}: {ms * 1e3:8.1f} us  "
                          f"{flops / ms / 1e9:7.1f} TFLOP/s  {byts / ms / 1e6:7.0f} GB/s", flush=True)
