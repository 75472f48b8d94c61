#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-s> [--gpus N] -- <command>; retries while the pod answers "transient / busy" (nothing charged)
T=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|nothing was charged\|no box or slot"; then
    echo "[retry $i] transient, sleeping 90 s"; sleep 90; continue
  fi
  break
done
