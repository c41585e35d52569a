M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
K='regex:conv_|stem_|head_|preprocess_'
timeout 900 ncu --metrics $M --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/step_metrics.csv python tools/profile_step.py 64 > gpurun_out/ncu_step.log 2>&1; echo "ncu step rc=$?"
WD_NO_SHIFT=1 timeout 900 ncu --metrics $M --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/step_metrics_noshift.csv python tools/profile_step.py 64 > gpurun_out/ncu_step_noshift.log 2>&1; echo "ncu noshift rc=$?"
python bench.py --workload videos --steps 3 --warmup 1 > gpurun_out/bench_videos32.json 2> gpurun_out/bench_videos32.err; echo "videos rc=$?"; cat gpurun_out/bench_videos32.json
python bench.py --arch tdn --batch 128 --num-class 11 --steps 10 --warmup 3 > gpurun_out/bench_tdn128.json 2> gpurun_out/bench_tdn128.err; echo "tdn rc=$?"; cat gpurun_out/bench_tdn128.json
python bench.py --num-class 11 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tsm11.json 2> gpurun_out/bench_tsm11.err; echo "tsm11 rc=$?"; cat gpurun_out/bench_tsm11.json
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_fuse2e|conv_fuse3|stem_pool2|preprocess_u8|conv_strip2_kernel|head_split" -s 20 -c 12 -o gpurun_out/r02_full_top python tools/profile_step.py 64 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
