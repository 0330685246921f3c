#!/bin/bash
# compute-sanitizer over the production kernels: tools/sanitize.sh TAG  -> gpurun_out/sanitize_TAG_{memcheck,racecheck,synccheck,initcheck}.log
TAG=${1:-x}
for tool in memcheck racecheck synccheck initcheck; do
  LOG=gpurun_out/sanitize_${TAG}_${tool}.log
  timeout -s KILL 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > $LOG 2>&1
  echo "$tool rc=$?" | tee -a $LOG
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target done|^ok " $LOG | tail -12
done
