#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_maingame.py shipped > gpurun_out/r02b_debug.txt 2>&1; echo "debug rc=$?"
tail -8 gpurun_out/r02b_debug.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -5 gpurun_out/r02b_pytest.log
timeout 300 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
tail -20 gpurun_out/r02b_bench.err
