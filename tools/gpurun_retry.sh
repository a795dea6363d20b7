#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <logfile> <command...>   -- retries while the pod answers "busy" (exit code 3)
# GPUS=N in the environment asks for N GPUs.
T=$1; LOG=$2; shift 2
for i in $(seq 1 30); do
  if [ -n "$GPUS" ]; then
    /usr/local/graft/bin/gpurun --gpus $GPUS --timeout $T -- "$@" > $LOG 2>&1
  else
    /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $LOG 2>&1
  fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
