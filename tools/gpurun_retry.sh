#!/bin/bash
# usage: tools/gpurun_retry.sh <gpurun args...>   (retries while the pod answers busy / transient: exit codes 2-3 are not charged)
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then
    sleep 150
    continue
  fi
  echo "$out" | tail -40
  exit $rc
done
echo "gpurun_retry: gave up"
exit 3
