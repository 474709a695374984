#!/bin/bash
# usage: tools/gpurun_retry.sh TAG TIMEOUT jobs...   retries while the pod answers busy (rc 3)
TAG=$1; TMO=$2; shift 2
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "bash tools/gpu_batch.sh $TAG $*" > gpurun_out/${TAG}_call.log 2>&1
  if grep -q "status=transient" gpurun_out/${TAG}_call.log; then sleep 150; continue; fi
  break
done
tail -8 gpurun_out/${TAG}_call.log
