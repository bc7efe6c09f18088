#!/bin/bash
# Run one gpurun call with retries on "transient / busy" answers (exit code 3 or status=transient); log to gpurun_out/call_<tag>.log
#   tools/gpucall.sh <tag> <timeout_s> <gpus> '<command>'
TAG=$1; TMO=$2; GPUS=$3; shift 3
LOG=gpurun_out/call_${TAG}.log
for try in 1 2 3 4 5 6 7 8 9 10; do
  if [ "$GPUS" = "1" ]; then
    /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  else
    /usr/local/graft/bin/gpurun --gpus $GPUS --timeout $TMO -- "$@" > $LOG 2>&1
  fi
  rc=$?
  if grep -q "status=transient\|status=busy" $LOG || [ $rc -eq 3 ]; then sleep 60; continue; fi
  break
done
echo "gpucall $TAG finished rc=$rc try=$try" >> $LOG
