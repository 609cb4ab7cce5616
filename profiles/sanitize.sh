#!/bin/bash
# compute-sanitizer over the mini-workload (run on a GPU box): profiles/sanitize.sh OUTDIR
out="${1:-gpurun_out/sanitize}"; mkdir -p "$out"
for tool in ${TOOLS:-memcheck racecheck synccheck initcheck}; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_driver.py 12 13 > "$out/$tool.log" 2>&1
  echo "== $tool rc=$? : $(grep -c 'ERROR SUMMARY' "$out/$tool.log") summary line(s): $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' "$out/$tool.log" | tail -1)"
done
