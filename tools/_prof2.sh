for env in "TVQ_STREAM_SETS=1" "TVQ_STREAM_SETS=2"; do
echo "=== $env"
for s in "1048576 512 64" "1048576 1024 128" "1048576 4096 128"; do
  env $env timeout 120 python tools/profile_stream.py run $s 1 2>&1 | grep -v Warn
done; done
