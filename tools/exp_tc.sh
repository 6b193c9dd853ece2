#!/bin/bash
# timing experiments for the tensor-core kernels: B2048_EXP is a comma list of -D macros
for e in "$@"; do
  touch 2048nn_b200/csrc/convnet_tc.cuh
  B2048_EXP=$e python 2048nn_b200/build.py >/dev/null 2>&1
  echo "== EXP [$e]"; timeout 100 python tools/tc_forward_bench.py 2>&1 | grep -E "fp16 n=(131072|1048576)"
  timeout 100 python tools/nn_playout_bench.py 2>&1 | grep C4
done
