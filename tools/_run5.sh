python -m pytest tests/test_gpu_round2.py -q -k "preprocess" 2>&1 | tail -5
python -m pytest tests/test_gpu_parity.py -q -k "preprocess or transform or golden" 2>&1 | tail -5
python tools/time_misc.py 2>&1 | tail -5
WD_PRE_PAIR=0 python tools/time_misc.py 2>&1 | head -1
