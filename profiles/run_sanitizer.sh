#!/bin/bash
# compute-sanitizer (ONE tool per call, B200_PROFILING.md) over the small end-to-end run, after the same run plain
cd /root/repo
TOOL=${1:-racecheck}
O=gpurun_out/r2san; mkdir -p $O
timeout 300 python profiles/small_end_to_end.py > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; exit 1; }
timeout 1200 compute-sanitizer --tool $TOOL --print-limit 20 python profiles/small_end_to_end.py > $O/$TOOL.log 2>&1; echo "sanitizer rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" $O/$TOOL.log | head -30
tail -5 $O/$TOOL.log
