#!/bin/bash
# A/B several builds of libgmdist on the same GPU box:
#   bash tools/ab.sh "<cases>" lib1.so lib2.so ...
CASES=$1; shift
for L in "$@"; do
  echo "=== $L"
  GMDIST_LIB=$PWD/$L python tools/dev_perf.py $CASES 2>&1 | grep -v "^NVIDIA"
done
