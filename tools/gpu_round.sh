#!/bin/bash
# One GPU-box pass.  Usage (repo root, under gpurun): bash tools/gpu_round.sh <tag> [steps...]
# steps: test perf bench ref launches full1m full1m_order fullcfg2
set -u
TAG=${1:-r1}; shift
STEPS=${*:-test perf bench ref}
OUT=gpurun_out
mkdir -p $OUT
for S in $STEPS; do
case $S in
test) timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu_$TAG.log;;
perf) timeout 600 python tools/dev_perf.py cfg2 cfg2split cfg3 cfg5s cfg5 > $OUT/dev_perf_$TAG.log 2>&1; echo "perf rc=$?"; cat $OUT/dev_perf_$TAG.log;;
bench) timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json
       timeout 600 python bench.py --workload cfg2_100k_x10_k16 --steps 20 --warmup 5 > $OUT/bench_cfg2_$TAG.json 2>> $OUT/bench_$TAG.err; echo "bench cfg2 rc=$?"; cat $OUT/bench_cfg2_$TAG.json;;
ref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err; echo "ref rc=$?"; cat $OUT/bench_ref_$TAG.json;;
launches) python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/plain_$TAG.log 2>&1 &&
   ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/ncu_launches_$TAG.log 2>&1; echo "launch list rc=$?";;
full1m) python tools/ncu_case.py 1000000 30 32 > $OUT/plain_case_$TAG.log 2>&1 &&
   ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 1 -c 1 -f -o $OUT/scan_1m_d30_$TAG \
    python tools/ncu_case.py 1000000 30 32 > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full 1M rc=$?";;
full1m_order) GMD_TC_ORDER=1 python tools/ncu_case.py 1000000 30 32 > $OUT/plain_case_order_$TAG.log 2>&1 &&
   GMD_TC_ORDER=1 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 1 -c 1 -f -o $OUT/scan_1m_d30_order_$TAG \
    python tools/ncu_case.py 1000000 30 32 > $OUT/ncu_full_order_$TAG.log 2>&1; echo "ncu full 1M ordered rc=$?";;
fullcfg2) python tools/ncu_case.py 100000 10 16 > $OUT/plain_case2_$TAG.log 2>&1 &&
   ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 1 -c 1 -f -o $OUT/scan_100k_d10_$TAG \
    python tools/ncu_case.py 100000 10 16 > $OUT/ncu_full2_$TAG.log 2>&1; echo "ncu full cfg2 rc=$?";;
esac
done
