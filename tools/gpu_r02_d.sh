#!/bin/bash
mkdir -p gpurun_out
for pct in 100 150 250; do B2048_MAIN_SINGLE_PCT=$pct timeout 100 python tools/debug_single_pct.py; done > gpurun_out/r02d_single_pct.txt 2>&1
cat gpurun_out/r02d_single_pct.txt
timeout 120 python tools/prof_main.py > gpurun_out/r02d_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:tc_persistent_kernel -s 1 -c 4 \
    -o gpurun_out/r02d_tc python tools/prof_main.py > gpurun_out/r02d_ncu_tc.log 2>&1
echo "ncu tc rc=$?"; tail -3 gpurun_out/r02d_ncu_tc.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:playout_fixed_kernel2 -c 1 \
    -o gpurun_out/r02d_k2 python tools/prof_main.py > gpurun_out/r02d_ncu_k2.log 2>&1
echo "ncu k2 rc=$?"
timeout 200 python bench.py --steps 2 > gpurun_out/r02d_bench_plain.json 2> gpurun_out/r02d_bench_plain.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02d_launches.csv \
    python bench.py --steps 2 > gpurun_out/r02d_bench_ncu.log 2>&1
echo "ncu launches rc=$?"
