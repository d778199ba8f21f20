#!/bin/bash
# gpurun with retries while the pod answers "transient" / busy (exit code 3, nothing charged)
# usage: tools/gpurun_retry.sh <timeout> <command...>
T=$1; shift
for attempt in $(seq 1 20); do
  out=$(gpurun --timeout "$T" "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|no box\|busy"; then
    echo "[retry $attempt] transient, sleeping 150 s"; sleep 150; continue
  fi
  echo "$out"; exit $rc
done
echo "gave up after 20 attempts"; exit 3
