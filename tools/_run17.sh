for v in 1 0 1 0; do echo "WD_F3_RES_PREFETCH=$v"; WD_F3_RES_PREFETCH=$v timeout 300 python tools/op_times.py 64 5 2>&1 | grep -E "layer2.[123].conv3|sum of"; done
timeout 600 python -m pytest tests/test_gpu_round2.py -q -x -k "fused_layer2" 2>&1 | tail -2
