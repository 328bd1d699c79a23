#!/bin/bash
# retry wrapper around gpurun: a busy pod answers rc 3 / "transient" without charging anything
# usage: scripts/gpu.sh [--gpus N] TIMEOUT 'command'
G=""
if [ "$1" == "--gpus" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun $G --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=3\|already running"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
