set -o pipefail
mkdir -p gpurun_out
timeout 200 python tools/prune_time.py 1000000 30 32 1.0 prune > gpurun_out/p2.log 2>&1 || { tail -5 gpurun_out/p2.log; exit 1; }
tail -1 gpurun_out/p2.log | cut -c1-150
timeout 100 python tools/prune_time.py 100000 10 16 1.0 plain prune > gpurun_out/p1.log 2>&1 || { tail -5 gpurun_out/p1.log; exit 1; }
tail -2 gpurun_out/p1.log | cut -c1-150
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log | grep -v Deprec
