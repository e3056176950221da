#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / status transient): usage
#   tools/gpurun_retry.sh <timeout_s> '<command>'
t=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then
    echo "attempt $attempt: busy, retrying in 90 s"; sleep 90; continue
  fi
  cat /tmp/gpurun_last.log | tail -60
  exit $rc
done
echo "gave up after 40 busy answers"; exit 3
