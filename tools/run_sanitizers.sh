#!/bin/bash
# compute-sanitizer over tools/san_case.py on the GPU box; summaries land in gpurun_out/ (copy them to profiles/).
#   bash tools/run_sanitizers.sh [per-tool timeout in seconds, default 900]
# racecheck only sees shared-memory hazards inside a CTA; the search kernel's tile-to-tile words are global memory
# written once per launch (tag + payload in one 64-bit store) and are covered by memcheck (addresses) and by the
# bit-exact parity tests (values).
T=${1:-900}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for tool in memcheck racecheck synccheck; do
  sections="host multi s4 pipe bands"
  # racecheck multiplies the run time of the shared-memory-heavy search by two orders of magnitude: the small cases
  [ "$tool" = racecheck ] && sections="host pipe"
  echo "== compute-sanitizer --tool $tool : $sections" > gpurun_out/sanitizer_$tool.txt
  timeout "$T" compute-sanitizer --tool $tool --print-limit 20 python tools/san_case.py $sections >> gpurun_out/sanitizer_$tool.txt 2>&1
  echo "exit code $?" >> gpurun_out/sanitizer_$tool.txt
  tail -5 gpurun_out/sanitizer_$tool.txt
done
