#!/bin/bash
# One GPU-box pass.  Usage (repo root, under gpurun): bash tools/gpu_round.sh <tag> [steps...]
# steps: test bench ref launches full1m full1m_dense fullcfg2 probes
set -u
TAG=${1:-r2}; shift
STEPS=${*:-test bench ref}
OUT=gpurun_out
mkdir -p $OUT
for S in $STEPS; do
case $S in
test) timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log | grep -v Deprec;;
bench) timeout 1200 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cut -c1-400 $OUT/bench_$TAG.json;;
ref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err; echo "ref rc=$?"; cut -c1-300 $OUT/bench_ref_$TAG.json;;
launches) python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --no-verify > $OUT/plain_$TAG.log 2>&1 &&
   ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --no-verify > $OUT/ncu_launches_$TAG.log 2>&1; echo "launch list rc=$?";;
full1m) python tools/ncu_case.py 1000000 30 32 > $OUT/plain_case_$TAG.log 2>&1 &&
   ncu --set full --clock-control none --import-source on -k regex:scan_tc_prune -s 1 -c 1 -f -o $OUT/scan_1m_d30_pruned_$TAG \
    python tools/ncu_case.py 1000000 30 32 > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full 1M pruned rc=$?";;
full1m_dense) GMD_TC_PRUNE=0 GMD_TC_ORDER=1 python tools/ncu_case.py 1000000 30 32 > $OUT/plain_case_dense_$TAG.log 2>&1 &&
   GMD_TC_PRUNE=0 GMD_TC_ORDER=1 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 1 -c 1 -f -o $OUT/scan_1m_d30_ordered_$TAG \
    python tools/ncu_case.py 1000000 30 32 > $OUT/ncu_full_dense_$TAG.log 2>&1; echo "ncu full 1M ordered full sweep rc=$?";;
fullcfg2) python tools/ncu_case.py 100000 10 16 > $OUT/plain_case2_$TAG.log 2>&1 &&
   ncu --set full --clock-control none --import-source on -k regex:scan_tc_prune -s 1 -c 1 -f -o $OUT/scan_100k_d10_pruned_$TAG \
    python tools/ncu_case.py 100000 10 16 > $OUT/ncu_full2_$TAG.log 2>&1; echo "ncu full cfg2 rc=$?";;
refine) python tools/ncu_case.py 1000000 30 32 > /dev/null 2>&1 &&
   ncu --set full --clock-control none --import-source on -k regex:refine_kernel -s 1 -c 1 -f -o $OUT/refine_1m_$TAG \
    python tools/ncu_case.py 1000000 30 32 > $OUT/ncu_refine_$TAG.log 2>&1; echo "ncu full refine rc=$?";;
esac
done
