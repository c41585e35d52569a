timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c.json 2>gpurun_out/bench_c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c.json')); print('b64', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])"
WD_STRIP7=0 WD_PRE_PAIR=0 WD_TAIL_SPLIT=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c0.json 2>gpurun_out/bench_c0.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c0.json')); print('b64 old', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])"
