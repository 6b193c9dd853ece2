#!/bin/bash
mkdir -p gpurun_out
for l in A B C D E F; do echo "== lib$l"; B2048_LIB=$PWD/2048nn_b200/exp/lib$l.so timeout 120 python tools/quick_k2.py 2 | tail -2; done 2>&1 | tee gpurun_out/r02k_k2.txt
for v in 1 2; do echo "== variant $v"; timeout 200 python tools/e2e_chunks.py $v; done 2>&1 | tee gpurun_out/r02k_e2e_chunks.txt
