#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/prof_k2_full.py > gpurun_out/r02q_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:playout_fixed_kernel3 -s 1 -c 1 \
    -o gpurun_out/r02q_k2v3 python tools/prof_k2_full.py > gpurun_out/r02q_ncu_k2.log 2>&1
echo "ncu k2 rc=$?"
