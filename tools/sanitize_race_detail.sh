#!/bin/bash
# racecheck with full hazard records (tools/sanitize.sh prints summaries only); expects /tmp/san.py from sanitize.sh
compute-sanitizer --tool racecheck --racecheck-report all --print-limit 40 python /tmp/san.py 2>&1 | grep -v "^c2_\|^c4_\|^mano" | head -150
