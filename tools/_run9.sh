WD_2CTA_TAP_MIN_KB=6 timeout 300 python tools/op_times.py 64 5 2>&1 | grep -E "layer2.0|sum of|forward"
timeout 300 python tools/op_times.py 64 5 2>&1 | grep -E "layer2.0|sum of|forward"
