#!/bin/bash
# usage: tools/gpurun_retry_n.sh NGPUS TAG TIMEOUT jobs...
N=$1; TAG=$2; TMO=$3; shift 3
for attempt in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --gpus $N --timeout $TMO -- "bash tools/gpu_batch.sh $TAG $*" > gpurun_out/${TAG}_call.log 2>&1
  if grep -q "status=transient" gpurun_out/${TAG}_call.log; then sleep 150; continue; fi
  break
done
tail -8 gpurun_out/${TAG}_call.log
