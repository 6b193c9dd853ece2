touch 2048nn_b200/csrc/convnet_tc.cuh
B2048_EXP=B2048_TRACE=3 python 2048nn_b200/build.py > /dev/null 2>&1 || python 2048nn_b200/build.py 2>&1 | grep error
timeout 60 python tools/tc_trace.py
