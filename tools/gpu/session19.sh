#!/bin/bash
# Round-2 GPU session 19 (1 GPU): compute-sanitizer (memcheck, racecheck, initcheck, synccheck) over every kernel family incl. the
# guessed select and the certified rounds (tools/sanitize_run.py).
set -x
O=gpurun_out/s19; mkdir -p $O
for tool in memcheck racecheck synccheck initcheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 1 python tools/sanitize_run.py > $O/sanitizer_$tool.txt 2>&1; echo "exit code $?" >> $O/sanitizer_$tool.txt
  tail -4 $O/sanitizer_$tool.txt
done
ls -la $O
