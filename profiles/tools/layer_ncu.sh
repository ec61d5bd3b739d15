#!/bin/bash
# ncu --set full of the kernels around the stage loop (final build), 250x250 elements, nop 4, 3 layers
cd $GRAFT_REPO_ROOT
ncu --clock-control none --set full -k regex:'k_(mom|layer|cons|bcl|pair_pack|btp_fin)' -c 9 -o /tmp/prof_layer -f python bench.py --steps 1 --warmup 0 --nelx 250 --nely 250 --no-cpu-baseline --no-e2e --no-config5 --no-cross-check > gpurun_out/layer_ncu.log 2>&1
ncu -i /tmp/prof_layer.ncu-rep --page raw --csv > gpurun_out/prof_layer_final.raw.csv 2>/dev/null
ls -la gpurun_out/prof_layer_final.raw.csv; tail -2 gpurun_out/layer_ncu.log
