#!/bin/bash
for e in "$@"; do
  touch 2048nn_b200/csrc/engine_kernels.cuh
  B2048_EXP=$e python 2048nn_b200/build.py >/dev/null 2>&1
  echo "== EXP [$e]"; timeout 100 python tools/quick_k2.py 2>&1 | grep "variant=1 n=16777216"
done
