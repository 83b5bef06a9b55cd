#!/bin/bash
# bash tools/ab_scan.sh "<scan_only args>;<scan_only args>..." lib1 lib2 ...
IFS=';' read -ra CASES <<< "$1"; shift
for L in "$@"; do
  echo "=== $L"
  for C in "${CASES[@]}"; do GMDIST_LIB=$PWD/$L python tools/scan_only.py $C 2>&1 | tail -1; done
done
