set -o pipefail
mkdir -p gpurun_out
timeout 60 python tools/prune_time.py 6000 10 16 1.0 plain prune > gpurun_out/p0.log 2>&1 || { tail -8 gpurun_out/p0.log; exit 1; }
tail -2 gpurun_out/p0.log | cut -c1-150
timeout 120 python tools/prune_time.py 100000 10 16 1.0 plain prune > gpurun_out/p1.log 2>&1 || { tail -5 gpurun_out/p1.log; exit 1; }
tail -2 gpurun_out/p1.log | cut -c1-150
timeout 200 python tools/prune_time.py 1000000 30 32 1.0 order prune > gpurun_out/p2.log 2>&1 || { tail -5 gpurun_out/p2.log; exit 1; }
tail -2 gpurun_out/p2.log | cut -c1-150
timeout 900 python -m pytest tests/test_prune_gpu.py -x -q -m gpu > gpurun_out/t1.log 2>&1; tail -4 gpurun_out/t1.log | grep -v Deprec
