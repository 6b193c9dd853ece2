#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
tail -25 gpurun_out/r02e_pytest.log
timeout 300 python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"
tail -4 gpurun_out/r02e_bench.err
python __graft_entry__.py smoke > gpurun_out/r02e_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02e_smoke.txt
