L=t-vq-vae-trajgen_b200/libtvq_b200.so
cp $L /tmp/cur.so
echo "== cur (sleep 64 after 8)"; timeout 300 python tools/time_sweep2.py 2>&1 | tail -9
for v in a b c d; do
cp tools/_lib_$v.so $L
echo "== $v"; timeout 300 python tools/time_sweep2.py 2>&1 | tail -9
done
cp /tmp/cur.so $L
