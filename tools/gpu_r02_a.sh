#!/bin/bash
# round-2 GPU pass A: tests, default bench, sweeps (each bounded by its own timeout)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 420 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02a_bench.err
timeout 300 python tools/sweeps.py > gpurun_out/r02a_sweeps.txt 2>&1; echo "sweeps rc=$?"
