#!/bin/bash
# compute-sanitizer pass over every kernel of libhso_b200.so (run on the GPU box: gpurun -- bash tools/sanitize.sh <tag>).
# Summaries land in gpurun_out/<tag>_sanitizer_<tool>_<part>.txt; copy the ones to keep into profiles/.
tag=${1:-r2}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for part in track frame match; do
    out=gpurun_out/${tag}_sanitizer_${tool}_${part}.txt
    timeout 900 $CS --tool $tool --print-limit 30 --error-exitcode 9 python tools/sanitize_workload.py $part > $out 2>&1
    echo "exit $?" >> $out
    echo "== $tool $part: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|exit ' $out | tr '\n' ' ')"
  done
done
