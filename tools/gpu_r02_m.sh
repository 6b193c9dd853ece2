#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout 150 -x > gpurun_out/r02m_pytest.log 2>&1; echo "pytest default rc=$?" >> gpurun_out/r02m_pytest.log
B2048_LIB=$PWD/2048nn_b200/exp/libE.so timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout 150 -x >> gpurun_out/r02m_pytest.log 2>&1; echo "pytest pool rc=$?" >> gpurun_out/r02m_pytest.log
grep -a "passed\|failed\|rc=" gpurun_out/r02m_pytest.log
for l in A B C D E F G H; do echo "== lib$l"; B2048_LIB=$PWD/2048nn_b200/exp/lib$l.so timeout 120 python tools/quick_k2.py 2; done 2>&1 | tee gpurun_out/r02m_k2.txt
