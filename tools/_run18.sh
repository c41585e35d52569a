python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2>gpurun_out/bench_final.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print('b64', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['blocking_call_value'], d['clocks'], d['cpu_baseline']['value'])"
sleep 20
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; echo "ref rc=$?"; head -c 600 gpurun_out/bench_ref.json; echo
timeout 300 python tools/op_times.py 64 5 > gpurun_out/op_times_final.log 2>&1; tail -4 gpurun_out/op_times_final.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --metrics $M --clock-control none -k "regex:conv_|stem_|head_|preprocess_" -c 200 --csv --log-file gpurun_out/step_metrics_final.csv python tools/profile_step.py 64 > gpurun_out/ncu_step_final.log 2>&1; echo "ncu step rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
