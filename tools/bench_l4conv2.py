import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from workoutdetector_b200.engine import bench_conv
for clips in (64, 63, 8):
    ms = bench_conv(clips, 7, 512, 512, 3, 1, 0, False, "tap", 256, 3, 20)
    fl = 2.0 * clips * 8 * 49 * 512 * 512 * 9
    print(f"WD_STRIP7={os.environ.get('WD_STRIP7','1')} l4.conv2 clips {clips}: {ms*1e3:.1f} us {fl/ms/1e9:.0f} TFLOP/s", flush=True)
