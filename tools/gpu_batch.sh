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
    bench)     timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ;;
    bench_noqr) B200_SVD_QR=0 timeout 1500 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_noqr.json 2> gpurun_out/${TAG}_bench_noqr.err ;;
    bench_nopred) B200_SVD_PREDICT=0 timeout 1500 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_nopred.json 2> gpurun_out/${TAG}_bench_nopred.err ;;
    bench_x32) B200_SVD_XROWS=32 B200_SVD_WROWS=48 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_x32.json 2> gpurun_out/${TAG}_bench_x32.err ;;
    bench_x48) B200_SVD_XROWS=48 B200_SVD_WROWS=64 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_x48.json 2> gpurun_out/${TAG}_bench_x48.err ;;
    bench_x16) B200_SVD_XROWS=16 B200_SVD_WROWS=32 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_x16.json 2> gpurun_out/${TAG}_bench_x16.err ;;
    tebd0)     timeout 900 python tools/bench_rows.py tebd > gpurun_out/${TAG}_tebd0.jsonl 2> gpurun_out/${TAG}_tebd0.err ;;
    tebd1)     B200_SVD_QR_COSTOL=1 timeout 900 python tools/bench_rows.py tebd > gpurun_out/${TAG}_tebd1.jsonl 2> gpurun_out/${TAG}_tebd1.err ;;
    tebd2)     B200_SVD_QR_COSTOL=1 B200_SVD_QR_TALL=16 timeout 900 python tools/bench_rows.py tebd > gpurun_out/${TAG}_tebd2.jsonl 2> gpurun_out/${TAG}_tebd2.err ;;
    tebdp4)    TEBD_PARALLEL=4 timeout 900 python tools/bench_rows.py tebd > gpurun_out/${TAG}_tebdp4.jsonl 2> gpurun_out/${TAG}_tebdp4.err ;;
    tebdp8)    TEBD_PARALLEL=8 timeout 900 python tools/bench_rows.py tebd > gpurun_out/${TAG}_tebdp8.jsonl 2> gpurun_out/${TAG}_tebdp8.err ;;
    tebd3)     B200_SVD_QR_COSTOL=1 B200_SVD_QR_TALL=16 B200_SVD_QR_MINQ=128 timeout 900 python tools/bench_rows.py tebd > gpurun_out/${TAG}_tebd3.jsonl 2> gpurun_out/${TAG}_tebd3.err ;;
    tebd4)     B200_SVD_QR_COSTOL=1 B200_SVD_QR_TALL=16 B200_SVD_QR_MINQ=64 timeout 900 python tools/bench_rows.py tebd > gpurun_out/${TAG}_tebd4.jsonl 2> gpurun_out/${TAG}_tebd4.err ;;
    tebdtest2) B200_SVD_QR_COSTOL=1 B200_SVD_QR_TALL=16 timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k tebd > gpurun_out/${TAG}_tebdtest2.log 2>&1 ;;
    smoke)     timeout 600 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${TAG}_smoke.log 2>&1 ;;
    rows)      timeout 1500 python tools/bench_rows.py > gpurun_out/${TAG}_rows.jsonl 2> gpurun_out/${TAG}_rows.err ;;
    svdstep)   timeout 900 python tools/svd_profile_step.py 40 > gpurun_out/${TAG}_svdstep.log 2>&1 ;;
    stepprof)  timeout 600 python tools/step_profile.py 50 300 > gpurun_out/${TAG}_stepprof.jsonl 2> gpurun_out/${TAG}_stepprof.err ;;
    stepprof_noqr) B200_SVD_QR=0 timeout 600 python tools/step_profile.py 50 300 > gpurun_out/${TAG}_stepprof_noqr.jsonl 2> gpurun_out/${TAG}_stepprof_noqr.err ;;
    phases)    B200_SVD_PHASES=1 timeout 900 python tools/qr_check.py oracle25 > gpurun_out/${TAG}_phases.log 2>&1 ;;
    ph00)      B200_SVD_QR_FASTP=0 B200_SVD_QR_FUSED=0 B200_SVD_PHASES=1 timeout 900 python tools/qr_check.py oracle25 > gpurun_out/${TAG}_ph00.log 2>&1 ;;
    ph10)      B200_SVD_QR_FASTP=768 B200_SVD_QR_FUSED=0 B200_SVD_PHASES=1 timeout 900 python tools/qr_check.py oracle25 > gpurun_out/${TAG}_ph10.log 2>&1 ;;
    ph01)      B200_SVD_QR_FASTP=0 B200_SVD_QR_FUSED=1 B200_SVD_PHASES=1 timeout 900 python tools/qr_check.py oracle25 > gpurun_out/${TAG}_ph01.log 2>&1 ;;
    bench_ref) timeout 1500 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err ;;
    bench_old) timeout 900 python bench.py --steps 20 --warmup 5 --preroll 0 > gpurun_out/${TAG}_bench_old.json 2> gpurun_out/${TAG}_bench_old.err ;;
    teacher)   timeout 1500 python tools/teacher_forced.py 7 26 > gpurun_out/${TAG}_teacher_7_26.json 2> gpurun_out/${TAG}_teacher.err ;;
    teacher2)  timeout 1500 python tools/teacher_forced.py 32 34 > gpurun_out/${TAG}_teacher_32_34.json 2> gpurun_out/${TAG}_teacher2.err ;;
    batch)     timeout 900 python -m pytest tests/test_batch_gpu.py -x -q > gpurun_out/${TAG}_batch.log 2>&1 ;;
    batchbench) timeout 900 python tools/batch_bench.py 592 60 25 > gpurun_out/${TAG}_batchbench.json 2> gpurun_out/${TAG}_batchbench.err ;;
    ncu_svd)   timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"qrcp_kernel|jacobi_kernel|apply_q_kernel" -c 6 -f -o gpurun_out/${TAG}_svd python tools/ncu_targets.py svd > gpurun_out/${TAG}_ncu_svd.log 2>&1 ;;
    ncu_batch) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tempo_batch_step_kernel -s 22 -c 1 -f -o gpurun_out/${TAG}_batch python tools/ncu_targets.py batch > gpurun_out/${TAG}_ncu_batch.log 2>&1 ;;
    ncu_win)   B200_BENCH_CUPROF=1 timeout 2400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv --log-file gpurun_out/${TAG}_launches_win.csv python bench.py --steps 2 --warmup 5 --preroll 25 --no-cpu > gpurun_out/${TAG}_ncu_win.log 2>&1 ;;
    ncu_list)  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 43000 -c 3300 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_list.log 2>&1 ;;
    c5)        timeout 1500 python tools/c5_bench.py 64 200 > gpurun_out/${TAG}_c5_n1.json 2> gpurun_out/${TAG}_c5_n1.err ;;
    c5r8on)    timeout 900 python tools/c5_bench.py 23 200 > gpurun_out/${TAG}_c5r8on.json 2> gpurun_out/${TAG}_c5r8on.err ;;
    c5r8off)   B200_BATCH_REBALANCE=0 timeout 900 python tools/c5_bench.py 23 200 > gpurun_out/${TAG}_c5r8off.json 2> gpurun_out/${TAG}_c5r8off.err ;;
    c5m_res)   timeout 900 python tools/c5_bench.py 32 200 > gpurun_out/${TAG}_c5m_res.json 2> gpurun_out/${TAG}_c5m_res.err ;;
    c5m_nores) B200_BATCH_RESERVE=0 timeout 900 python tools/c5_bench.py 32 200 > gpurun_out/${TAG}_c5m_nores.json 2> gpurun_out/${TAG}_c5m_nores.err ;;
    c5probe)   timeout 600 python tools/c5_overflow_probe.py 32 200 32 > gpurun_out/${TAG}_c5probe.json 2> gpurun_out/${TAG}_c5probe.err ;;
    c5small)   timeout 900 python tools/c5_bench.py 24 100 > gpurun_out/${TAG}_c5small.json 2> gpurun_out/${TAG}_c5small.err ;;
    c5_n2)     timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/c5_bench.py 64 200 > gpurun_out/${TAG}_c5_n2.json 2> gpurun_out/${TAG}_c5_n2.err ;;
    bench_n8)  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n8.json 2> gpurun_out/${TAG}_bench_n8.err ;;
    bench_n4)  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n4.json 2> gpurun_out/${TAG}_bench_n4.err ;;
    c5_n4)     timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 tools/c5_bench.py 64 200 > gpurun_out/${TAG}_c5_n4.json 2> gpurun_out/${TAG}_c5_n4.err ;;
    c5_n8r)    B200_BATCH_RESERVE=24 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 tools/c5_bench.py 64 200 > gpurun_out/${TAG}_c5_n8r.json 2> gpurun_out/${TAG}_c5_n8r.err ;;
    c5_n8)     timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/c5_bench.py 64 200 > gpurun_out/${TAG}_c5_n8.json 2> gpurun_out/${TAG}_c5_n8.err ;;
    bench_n2)  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err ;;
    *) echo "unknown job $job" ;;
  esac
  echo "== $job rc=$?"
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
