L=t-vq-vae-trajgen_b200/libtvq_b200.so
cp $L /tmp/cur.so
for i in 1 2; do
echo "== cur"; timeout 300 python tools/time_sweep2.py small 2>&1 | tail -4
cp tools/_libprev.so $L
echo "== prev"; timeout 300 python tools/time_sweep2.py small 2>&1 | tail -4
cp /tmp/cur.so $L
done
