#!/bin/bash
# bash tools/ab_tc.sh lib1 lib2 ... : tensor-core and FP32 scans of several builds on one box
for L in "$@"; do
  echo "=== $L"
  GMDIST_LIB=$PWD/$L python tools/tc_time.py 1000000 30 32 2>&1 | grep "dbg=0"
  GMDIST_LIB=$PWD/$L python tools/tc_time.py 100000 10 16 2>&1 | grep "dbg=0"
  GMDIST_LIB=$PWD/$L GMD_TC=0 python tools/dev_perf.py cfg2 cfg5s 2>&1 | grep cfg | sed 's/form=. kk=[0-9]* //; s/col_mean.*//'
done
