set -x
python bench.py > gpurun_out/r02s4_bench_1gpu.json 2> gpurun_out/r02s4_bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02s4_bench_reference_arm.json 2> gpurun_out/ref.err
python tools/op_times.py 64 5 > gpurun_out/r02s4_op_times.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s4_smoke.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02s4_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k 'regex:conv_|stem_|head_|preprocess_' -c 200 --csv --log-file gpurun_out/step_metrics_s4.csv python tools/profile_step.py 64 > gpurun_out/ncu_step.log 2>&1
ncu --set full --clock-control none -k 'regex:conv_|stem_|head_|preprocess_' -s 45 -c 45 -o /tmp/full_s4 -f python tools/profile_step.py 64 > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/full_s4.ncu-rep > gpurun_out/r02s4_ncu_full_step.txt 2>&1
ls -la /tmp/full_s4.ncu-rep
tail -2 gpurun_out/r02s4_op_times.log; tail -1 gpurun_out/r02s4_smoke.log; cut -c1-200 gpurun_out/r02s4_bench_1gpu.json
