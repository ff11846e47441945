for s in "1048576 512 64" "1048576 1024 128" "1048576 4096 128"; do
  timeout 120 python tools/profile_stream.py run $s 1 2>&1 | grep -E "n=|convert"
done
