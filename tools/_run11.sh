timeout 900 python -m pytest tests/test_gpu_tdn.py tests/test_gpu_round2.py -x -q 2>&1 | tail -4
timeout 600 python tools/op_times_tdn.py 128 3 > gpurun_out/op_times_tdn2.log 2>&1; grep -E "layer2|layer3.0|by kind|sum of|wd_forward" gpurun_out/op_times_tdn2.log
