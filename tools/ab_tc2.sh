#!/bin/bash
# bash tools/ab_tc2.sh "<dbg list>" lib1 lib2 ... : tensor-core scan timing of several builds on one box (interleaved twice)
DBG=$1; shift
for R in 1 2; do
for L in "$@"; do
  echo "=== $L"
  GMDIST_LIB=$PWD/$L timeout 200 python tools/tc_time.py 1000000 30 32 $DBG 2>&1 | grep "dbg="
  GMDIST_LIB=$PWD/$L timeout 120 python tools/tc_time.py 100000 10 16 0 2>&1 | grep "dbg=0"
done
done
