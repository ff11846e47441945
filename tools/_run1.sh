timeout 600 python tools/check_stream.py 2>&1 | grep -v Warn | grep -E "FAIL|PARITY" | tail -5
timeout 300 python tools/time_sweep2.py 2>&1 | tail -12
