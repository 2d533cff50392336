#!/bin/bash
# Developer A/B harness: CUDA-event stage times of several builds of the library in ONE gpurun call (boxes differ by a
# few percent, so variants are only comparable inside one call).  usage: tools/ab.sh reps v0 v1 ...  (bayesbay_b200/ab/<v>.so)
reps=$1; shift
for r in $(seq $reps); do
  for v in "$@"; do
    echo -n "$v: "
    BBB_LIB_PATH=$PWD/bayesbay_b200/ab/$v.so timeout 200 python tools/stage_time.py 1024 300 50 2>&1 | tr '\n' ' '
    echo
  done
done
