#!/bin/bash
# A/B of environment switches on one library build: tools/ab_env.sh reps "VAR=a" "VAR=b" ...
reps=$1; shift
for r in $(seq $reps); do
  for v in "$@"; do
    echo -n "$v: "
    env $v timeout 200 python tools/stage_time.py 1024 300 50 2>&1 | tr '\n' ' '
    echo
  done
done
