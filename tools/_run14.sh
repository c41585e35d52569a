timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/op_times.py 64 5 > gpurun_out/op_times3.log 2>&1; tail -50 gpurun_out/op_times3.log
