#!/bin/bash
# bash tools/ab_tc3.sh "<dbg list>" lib1 lib2 ... : 1M x 30 k=32 and 100k x 10 k=16 tensor-core scan times of several builds
DBG=$1; shift
for L in "$@"; do
  echo "=== $L"
  GMDIST_LIB=$PWD/$L timeout 100 python tools/tc_time.py 1000000 30 32 $DBG 2>&1 | grep "dbg="
  GMDIST_LIB=$PWD/$L timeout 60 python tools/tc_time.py 100000 10 16 0 2>&1 | grep "dbg=0"
done
