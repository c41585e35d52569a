for i in 1 2; do timeout 300 python tools/op_times.py 64 5 2>&1 | grep -E "^head|sum of|forward"; done
