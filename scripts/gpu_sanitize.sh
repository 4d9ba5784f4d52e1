#!/bin/bash
# compute-sanitizer over small cases of every kernel variant (scripts/sanitize_case.py); logs go to gpurun_out/ and are
# copied to profiles/r2/ by hand.  $1 = tools to run (default "memcheck racecheck")
mkdir -p gpurun_out
for tool in ${1:-memcheck racecheck}; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/sanitizer_$tool.log \
      python scripts/sanitize_case.py > gpurun_out/sanitizer_$tool.out 2>&1
  echo "compute-sanitizer $tool exit $?"
  tail -3 gpurun_out/sanitizer_$tool.out
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|========= (Invalid|Race|Error)" gpurun_out/sanitizer_$tool.log | sort | uniq -c | head -12
done
