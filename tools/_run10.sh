timeout 600 python -m pytest tests -m gpu -x -q -k "host or stream or blocking or infer" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_d.json 2>gpurun_out/bench_d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_d.json')); print('b64', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['blocking_call_value'], d['e2e']['h2d_probe_gbs_per_gpu'], d['clocks'])"
timeout 600 python tools/op_times_tdn.py 128 3 > gpurun_out/op_times_tdn.log 2>&1; tail -100 gpurun_out/op_times_tdn.log
