#!/bin/bash
# compute-sanitizer over every kernel family (race / sync / memory checks); logs under gpurun_out/.
# Each tool runs under its own timeout so that a hung tool cannot take the box with it.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck initcheck; do
  echo "== $tool" > gpurun_out/sanitizer_$tool.log
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_probe.py >> gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit=$?" >> gpurun_out/sanitizer_$tool.log
  tail -4 gpurun_out/sanitizer_$tool.log
done
