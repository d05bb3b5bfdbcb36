#!/bin/bash
# usage: tools/gpurun_retry.sh <gpurun args...>  -- retries while the pod answers "transient" (nothing charged)
for attempt in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -n 60
  if echo "$out" | grep -q "status=transient"; then
    echo "[retry] attempt $attempt transient; sleeping 120 s"
    sleep 120
  else
    exit 0
  fi
done
