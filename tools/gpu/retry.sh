#!/bin/bash
# usage: tools/gpu/retry.sh <timeout> <script> [gpurun extra args...]: retries while the pod answers busy (exit code 3), up to 12 times
T=$1; S=$2; shift 2
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" --timeout $T -- "bash $S"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] pod busy, attempt $i; sleeping 150 s"; sleep 150
done
exit 3
