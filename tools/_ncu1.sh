for cfg in "s512 1048576 512 64 ILi64ELi256ELb1ELi2ELi4ELi1E" "s4096 1048576 4096 128 ILi128ELi256ELb1ELi2ELi4ELi1E"; do
  set -- $cfg
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:fwd_stream -s 2 -c 1 -f -o gpurun_out/$1 python tools/ncu_stream.py $2 $3 $4 > gpurun_out/$1.ncu.log 2>&1
  bash tools/ncu_digest.sh $1 $5
done
