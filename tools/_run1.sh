set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"
cat gpurun_out/bench_a.json
python tools/op_times.py 64 5 > gpurun_out/op_times.log 2>&1; echo "op_times rc=$?"
tail -70 gpurun_out/op_times.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -c 300 --csv --log-file gpurun_out/step_metrics.csv python tools/profile_step.py 64 > gpurun_out/ncu_step.log 2>&1; echo "ncu rc=$?"
