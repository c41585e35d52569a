N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_clips_${N}gpu.json 2> gpurun_out/bench_clips_${N}gpu.err; echo "clips rc=$?"; python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_clips_${N}gpu.json') if l.startswith('{')][-1])
print($N, 'clips', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['blocking_call_value'], d['e2e']['h2d_probe_gbs_per_gpu'], d['clocks'])
PY
