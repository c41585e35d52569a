timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "test_conv_umma_vs_torch" 2>&1 | tail -5
WD_STRIP7=0 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "test_conv_umma_vs_torch" 2>&1 | tail -2
timeout 120 python tools/bench_l4conv2.py
WD_STRIP7=0 timeout 120 python tools/bench_l4conv2.py
