#!/bin/bash
# same-box A/B of library builds: ab.sh <nel> <repeats> libA.so libB.so ...   (boxes differ by +-6 % in granted clocks: never compare across calls)
cd $GRAFT_REPO_ROOT
N=$1; R=$2; shift 2
O=gpurun_out/ab.log; : > $O
for r in $(seq 1 $R); do for lib in "$@"; do
  HNUMO_LIB_PATH=$PWD/h-numo_b200/$lib python profiles/tools/stage_time.py $N $ABOPTS >> $O 2>&1
done; done
cat $O
