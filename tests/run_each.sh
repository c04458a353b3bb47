#!/bin/bash
# Run every GPU test in its own process with a wall-clock limit, so a hung kernel costs one test, not the call.
# usage: tests/run_each.sh <file> <seconds-per-test>
f=$1; lim=${2:-120}
mkdir -p gpurun_out
: > gpurun_out/each.log
for t in $(python -m pytest "$f" -m gpu --collect-only -q 2>/dev/null | grep "::"); do
  s=$(date +%s)
  timeout -s KILL "$lim" python -m pytest "$t" -m gpu -x -q > gpurun_out/one.log 2>&1; rc=$?
  e=$(date +%s)
  echo "$t rc=$rc $((e - s))s" | tee -a gpurun_out/each.log
  if [ $rc -ne 0 ]; then tail -25 gpurun_out/one.log | tee -a gpurun_out/each.log; fi
done
