#!/bin/bash
# clock64 timeline of one forward of CTA 0: builds a -DB2048_TRACE=<iteration> variant next to the product library
export B2048_LIB=$PWD/2048nn_b200/lib2048nn_b200_trace.so
B2048_EXP=B2048_TRACE=${1:-3} python 2048nn_b200/build.py > /dev/null 2>&1 || { echo "trace build failed"; exit 1; }
timeout 60 python tools/tc_trace.py
