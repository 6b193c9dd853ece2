#!/bin/bash
# round-2 closing evidence at HEAD: probe test, sweeps (isolated + streamed), ncu of the tensor-core kernels, bench launch list
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -k "pipe_rate or variants" 2>&1 | tail -3
timeout 300 python tools/sweeps.py > gpurun_out/r02u_sweeps.txt 2>&1; echo "sweeps rc=$?"; cat gpurun_out/r02u_sweeps.txt
timeout 120 python tools/prof_main.py > gpurun_out/r02u_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:tc_persistent_kernel -s 1 -c 4 \
    -o gpurun_out/r02u_tc python tools/prof_main.py > gpurun_out/r02u_ncu_tc.log 2>&1
echo "ncu tc rc=$?"; tail -3 gpurun_out/r02u_ncu_tc.log
timeout 200 python bench.py --steps 2 > gpurun_out/r02u_bench_plain.json 2> gpurun_out/r02u_bench_plain.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02u_launches.csv \
    python bench.py --steps 2 > gpurun_out/r02u_bench_ncu.log 2>&1
echo "ncu launches rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02u_bench_plain.json").read().strip().splitlines()[-1])
print(json.dumps(d["roofline"]["executed"], indent=1))
PY
