#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/debug_maingame.py shipped > gpurun_out/r02c_debug.txt 2>&1; echo "debug rc=$?"
tail -14 gpurun_out/r02c_debug.txt
timeout 700 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -15 gpurun_out/r02c_pytest.log
timeout 240 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
tail -6 gpurun_out/r02c_bench.err
