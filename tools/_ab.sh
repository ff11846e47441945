L=t-vq-vae-trajgen_b200/libtvq_b200.so
cp $L /tmp/cur.so
for i in 1 2; do
echo "== cur"; timeout 300 python tools/time_small.py 2>&1 | grep -E "^(18432|76800|307200)"
cp tools/_libprev.so $L
echo "== prev"; timeout 300 python tools/time_small.py 2>&1 | grep -E "^(18432|76800|307200)"
cp /tmp/cur.so $L
done
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
