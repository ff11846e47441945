for env in "TVQ_STREAM_SETS=2" "TVQ_STREAM_SETS=1"; do
echo "=== $env"; env $env timeout 300 python tools/time_sweep2.py 2>&1 | tail -9
done
