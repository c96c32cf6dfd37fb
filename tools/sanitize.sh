#!/bin/bash
# GPU box: compute-sanitizer (memcheck, racecheck, synccheck) over smoke() and a small search; logs -> gpurun_out/sanitize_*.log
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  for what in smoke search; do
    if [ "$what" = smoke ]; then cmd='python -c "import __graft_entry__ as g; g.smoke()"'; else cmd='python tools/small_search.py'; fi
    echo "=== $tool $what"
    timeout 900 bash -c "compute-sanitizer --tool $tool --print-limit 20 $cmd" > gpurun_out/sanitize_${tool}_${what}.log 2>&1
    echo "rc=$?" >> gpurun_out/sanitize_${tool}_${what}.log
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke OK|small search OK|rc=" gpurun_out/sanitize_${tool}_${what}.log | tail -4
  done
done
