#!/bin/bash
# usage: tools/gpu_retry.sh [--gpus N] <timeout_s> '<command>'   -- retries while the pod answers "busy" (nothing charged)
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun $GP --timeout $T -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then sleep 75; continue; fi
  echo "$OUT"; exit 0
done
echo "$OUT"; echo "gave up after 40 busy answers"; exit 3
