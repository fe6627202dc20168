#!/bin/bash
# synccheck (barrier misuse) and initcheck (reads of uninitialised device memory) over the same script as tools/sanitize.sh
# (run tools/sanitize.sh first: it writes /tmp/san.py)
for tool in synccheck initcheck; do
  echo "== compute-sanitizer --tool $tool"
  compute-sanitizer --tool $tool --print-limit 8 python /tmp/san.py 2>&1 | grep -E "ERROR SUMMARY|Error|Uninitialized|Barrier|at |c2_full|c4_arm|mano_fit" | grep -v "Host Frame" | head -40
done
