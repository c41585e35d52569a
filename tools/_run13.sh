timeout 120 python tools/bench_s2.py
WD_S2_PREFETCH=0 timeout 120 python tools/bench_s2.py
timeout 120 python tools/bench_s2.py
WD_S2_PREFETCH=0 timeout 120 python tools/bench_s2.py
