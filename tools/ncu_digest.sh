#!/bin/bash
# usage: ncu_digest.sh <name> <kernel-substring-for-ncu_lines> : summary + per-line attribution of gpurun_out/<name>.ncu-rep -> text, then drop the report
n=$1
python tools/ncu_summary.py gpurun_out/$n.ncu-rep > gpurun_out/$n.summary.txt 2>&1
python tools/ncu_lines.py gpurun_out/$n.ncu-rep t-vq-vae-trajgen_b200/libtvq_b200.so "$2" 40 > gpurun_out/$n.lines.txt 2>&1
BY_SAMPLES=1 python tools/ncu_lines.py gpurun_out/$n.ncu-rep t-vq-vae-trajgen_b200/libtvq_b200.so "$2" 25 > gpurun_out/$n.stalls.txt 2>&1
rm -f gpurun_out/$n.ncu-rep
