#!/bin/bash
# usage: scratch/gpu.sh <logfile> <gpus> <timeout> '<command>'   -- retries while the pod has no free slot (exit code 3)
log=$1; gpus=$2; to=$3; shift 3
for i in $(seq 1 20); do
  if [ "$gpus" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1; else /usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "$@" > $log 2>&1; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 75
done
exit 3
