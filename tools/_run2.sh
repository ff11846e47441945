for env in "TVQ_STREAM_SP=4" "TVQ_STREAM_SP=2" "TVQ_STREAM_CG=2" "TVQ_STREAM_CG=1" "TVQ_STREAM_XD=4" "TVQ_STREAM_XD=2"; do
echo "=== $env"; env $env timeout 300 python tools/time_sweep2.py 2>&1 | tail -9
done
