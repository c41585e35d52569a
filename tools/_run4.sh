python tools/exp_gaps.py 64 > gpurun_out/exp_gaps.log 2>&1; echo "gaps rc=$?"; cat gpurun_out/exp_gaps.log | tail -12
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_fuse2e|conv_fuse3|stem_pool2|preprocess_u8|conv_strip2_kernel|head_split" -s 11 -c 11 -o gpurun_out/r02_full_top python tools/profile_step.py 64 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
