import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from workoutdetector_b200.engine import bench_conv
for name, H, Cin, Cout in (("l3.0.conv2", 28, 256, 256), ("l4.0.conv2", 14, 512, 512)):
    ms = bench_conv(64, H, Cin, Cout, 3, 2, 0, False, "tap", 256, 3, 20)
    fl = 2.0 * 64 * 8 * (H // 2) ** 2 * Cout * Cin * 9
    print(f"WD_STRIP7={os.environ.get('WD_STRIP7','2')} {name} clips 64: {ms*1e3:.1f} us {fl/ms/1e9:.0f} TFLOP/s", flush=True)
