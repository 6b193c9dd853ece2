#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/debug_maingame.py shipped > gpurun_out/r02f_debug.txt 2>&1; echo "debug rc=$?"
head -14 gpurun_out/r02f_debug.txt
timeout 800 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
tail -12 gpurun_out/r02f_pytest.log
timeout 300 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02f_bench.err
