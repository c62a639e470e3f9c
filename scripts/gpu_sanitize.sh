#!/bin/bash
O=gpurun_out; mkdir -p $O
for tool in memcheck racecheck initcheck synccheck; do
  echo "== compute-sanitizer $tool"
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize.py > $O/san_$tool.txt 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|mismatches|Error|hazard" $O/san_$tool.txt | head -8
done
