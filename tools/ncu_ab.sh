#!/bin/bash
# ncu --set full of the scan kernel for several builds: bash tools/ncu_ab.sh "<ncu_case args>" tag lib1 lib2 ...
ARGS=$1; TAG=$2; shift 2
for L in "$@"; do
  N=$(basename $L .so)
  GMDIST_LIB=$PWD/$L python tools/ncu_case.py $ARGS > gpurun_out/plain_${TAG}_$N.log 2>&1 &&
  GMDIST_LIB=$PWD/$L ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 1 -c 1 -f -o gpurun_out/${TAG}_$N \
     python tools/ncu_case.py $ARGS > gpurun_out/ncu_${TAG}_$N.log 2>&1
  echo "$N rc=$?"
done
