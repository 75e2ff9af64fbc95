#!/bin/bash
# usage: scripts/gpu/run.sh <call-script> <timeout-s> [gpus]   (retries while the pod is busy: exit code 3)
S=$1; T=${2:-1500}; N=${3:-1}
for i in $(seq 1 30); do
  if [ "$N" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "bash $S"; else /usr/local/graft/bin/gpurun --gpus $N --timeout $T -- "bash $S"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
