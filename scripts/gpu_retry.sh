#!/bin/bash
# Developer helper: run a command on the GPU box through gpurun, retrying while the pod answers "busy" (exit 3).
# The built libraries travel with the snapshot; they are touched first so that sources edited after the local build do not
# trigger a rebuild on the box.     usage: scripts/gpu_retry.sh <timeout_s> '<command>' [gpus]
T=$1; CMD=$2; G=${3:-1}
PRE='touch blitzen_b200/*.so oracle/*.so 2>/dev/null; mkdir -p gpurun_out; '
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$PRE$CMD"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$PRE$CMD"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
