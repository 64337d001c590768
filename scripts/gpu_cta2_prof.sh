#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/prof_net.py > gpurun_out/prof_net_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_halo2' -s 2 -c 1 -f -o gpurun_out/prof_halo2 python scripts/prof_net.py > gpurun_out/prof_halo2.log 2>&1
echo "ncu exit=$? :: $(tail -n 1 gpurun_out/prof_halo2.log)"
ncu -i gpurun_out/prof_halo2.ncu-rep --page raw --csv > gpurun_out/prof_halo2_raw.csv 2>/dev/null; echo "raw exit=$?"
