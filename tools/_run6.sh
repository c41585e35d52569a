python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_b.json 2>gpurun_out/bench_b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b.json')); print('b64', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])"
for b in 128 148; do python bench.py --batch $b --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/bench_batch$b.json 2>gpurun_out/bench_batch$b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_batch$b.json')); print('b$b', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])"; done
