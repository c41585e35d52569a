for b in 2 4 8 16; do python tools/op_times.py $b 10 > gpurun_out/op_times_b$b.log 2>&1; echo "b=$b rc=$?"; tail -4 gpurun_out/op_times_b$b.log; done
