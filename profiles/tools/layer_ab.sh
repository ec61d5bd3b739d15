#!/bin/bash
# per-kernel times of one step (ncu launch list) for several library builds: layer_ab.sh lib1.so lib2.so ...
cd $GRAFT_REPO_ROOT
: > gpurun_out/layer_ab.log
for lib in "$@"; do
  HNUMO_LIB_PATH=$PWD/h-numo_b200/$lib ncu --clock-control none --metrics gpu__time_duration.sum -c 900 --csv --log-file gpurun_out/layer_$lib.csv python bench.py --nelx 250 --nely 250 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-config5 --no-cross-check > /dev/null 2>&1
  echo "== $lib" >> gpurun_out/layer_ab.log
  python profiles/summarize_launches.py gpurun_out/layer_$lib.csv | grep -v "stage_pair\|nodal_to_quad\|diag\|total" | head -9 >> gpurun_out/layer_ab.log
done
cat gpurun_out/layer_ab.log
