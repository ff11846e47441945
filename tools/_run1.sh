timeout 600 python tools/check_stream.py 2>&1 | grep -v Warn | grep -E "FAIL|PARITY|dup|n=300000" | tail -12
timeout 300 python tools/time_sweep2.py 2>&1 | tail -12
