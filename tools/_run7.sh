timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "test_conv_umma_vs_torch" 2>&1 | tail -5
WD_STRIP2=0 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "test_conv_umma_vs_torch" 2>&1 | tail -2
timeout 120 python tools/bench_l4conv2.py
WD_TAIL_SPLIT=0 timeout 120 python tools/bench_l4conv2.py
timeout 300 python tools/op_times.py 64 5 > gpurun_out/op_times2.log 2>&1; grep -E "conv2|sum of|forward" gpurun_out/op_times2.log
