#!/bin/bash
# usage: tools/gpurun_retry.sh <tag> <timeout_s> [--gpus N] -- <command>   (retries while the pod answers busy, rc 3)
tag=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > gpurun_out/${tag}_gpurun.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
