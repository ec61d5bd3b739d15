#!/bin/bash
# ab2.sh <nel> <reps> "lib[:opt=val,...]" ...   same-box A/B of library builds with run-time options
cd $GRAFT_REPO_ROOT
N=$1; R=$2; shift 2
O=gpurun_out/ab.log; : > $O
for r in $(seq 1 $R); do for spec in "$@"; do
  lib=${spec%%:*}; opts=""; [ "$spec" != "$lib" ] && opts=$(echo ${spec#*:} | tr ',' ' ')
  HNUMO_LIB_PATH=$PWD/h-numo_b200/$lib python profiles/tools/stage_time.py $N $opts >> $O 2>&1
done; done
cat $O
