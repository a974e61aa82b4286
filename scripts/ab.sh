#!/bin/bash
# same-box A/B of library builds: scripts/ab.sh <cfg> <steps> <rounds> libA.so libB.so ...   (graph-replay step time, alternating)
CFG=$1; STEPS=$2; ROUNDS=$3; shift 3
for r in $(seq 1 $ROUNDS); do
  for L in "$@"; do
    echo -n "$(basename $L): "; WN_LIB=$PWD/$L python scripts/steptime.py $CFG $STEPS 2>&1 | tail -1
  done
done
