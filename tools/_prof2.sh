for s in "1048576 512 64" "1048576 1024 128" "1048576 4096 128" "524288 16384 256"; do
  timeout 120 python tools/profile_stream.py run $s 1 2>&1 | grep -v Warn
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
timeout 300 python tools/time_sweep2.py small 2>&1 | tail -4
