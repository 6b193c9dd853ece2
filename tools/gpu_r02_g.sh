#!/bin/bash
mkdir -p gpurun_out
timeout 400 bash tools/debug_asserts_run.sh > gpurun_out/r02g_debug_asserts.txt 2>&1; echo "debug asserts rc=$?"
tail -8 gpurun_out/r02g_debug_asserts.txt
timeout 200 python tools/e2e_chunks.py > gpurun_out/r02g_e2e_chunks.txt 2>&1; echo "e2e rc=$?"; cat gpurun_out/r02g_e2e_chunks.txt
timeout 600 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log
tail -6 gpurun_out/r02g_pytest.log
