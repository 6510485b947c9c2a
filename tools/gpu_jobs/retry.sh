#!/bin/bash
# usage: retry.sh <timeout> <gpus> <script>   — retries while the pod answers "busy" (exit 3)
for i in $(seq 1 40); do
  if [ "$2" = "1" ]; then
    /usr/local/graft/bin/gpurun --timeout "$1" -- "bash $3"
  else
    /usr/local/graft/bin/gpurun --gpus "$2" --timeout "$1" -- "bash $3"
  fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
