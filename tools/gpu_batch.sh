#!/bin/bash
# One gpurun call = several bounded jobs, each with its own log under gpurun_out/.
# usage: tools/gpu_batch.sh TAG job1 job2 ...   (jobs: check tiny_san kernels parity bench bench_noqr rows)
TAG=$1; shift
mkdir -p gpurun_out
for job in "$@"; do
  case $job in
    check)     timeout 900 python tools/qr_check.py synthetic > gpurun_out/${TAG}_check.log 2>&1 ;;
    check25)   timeout 900 python tools/qr_check.py oracle25 > gpurun_out/${TAG}_check25.log 2>&1 ;;
    big)       timeout 900 python tools/qr_check.py big > gpurun_out/${TAG}_big.log 2>&1 ;;
    tiny_san)  timeout 600 compute-sanitizer --tool memcheck python tools/qr_check.py tiny > gpurun_out/${TAG}_memcheck.log 2>&1 ;;
    kernels)   timeout 1200 python -m pytest tests/test_kernels_gpu.py -x -q > gpurun_out/${TAG}_kernels.log 2>&1 ;;
    gputests)  timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputests.log 2>&1 ;;
    bench)     timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ;;
    bench_noqr) B200_SVD_QR=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_noqr.json 2> gpurun_out/${TAG}_bench_noqr.err ;;
    *) echo "unknown job $job" ;;
  esac
  echo "== $job rc=$?"
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
