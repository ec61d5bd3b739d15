#!/bin/bash
# Round-2 evidence run (one B200): driver-style bench line, launch list of the same command at 250x250, one `ncu --set full`
# capture of the stage kernel at nop 4 (500x500 elements) and nop 8 (200x200, 10 layers), one of the layer kernels.
# gpurun copies back at most 64 MiB: the reports are exported to CSV on the box and only the nop-4 report travels.
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2prof; mkdir -p $O
python bench.py --steps 3 --warmup 3 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
NCU="ncu --clock-control none"
B="python bench.py --steps 1 --no-cpu-baseline --no-e2e --no-config5 --no-cross-check"
$NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file $O/r2_launches_250x250.csv $B --nelx 250 --nely 250 --warmup 1 > $O/r2_ncu_l.log 2>&1
$NCU --set full --import-source on -k regex:k_btp_stage_pair -s 30 -c 1 -o $O/prof_r2_n4 -f $B --nelx 500 --nely 500 --warmup 0 > $O/r2_ncu_n4.log 2>&1
ncu -i $O/prof_r2_n4.ncu-rep --page raw --csv > $O/prof_r2_n4.raw.csv 2>/dev/null
ncu -i $O/prof_r2_n4.ncu-rep --page source --csv --print-source cuda,sass > $O/prof_r2_n4.source.csv 2>/dev/null
$NCU --set full --import-source on -k regex:k_btp_stage_pair -s 30 -c 1 -o /tmp/prof_r2_n8 -f $B --nop 8 --layers 10 --nelx 200 --nely 200 --warmup 0 > $O/r2_ncu_n8.log 2>&1
ncu -i /tmp/prof_r2_n8.ncu-rep --page raw --csv > $O/prof_r2_n8.raw.csv 2>/dev/null
$NCU --set full -k regex:'k_(mom|layer|cons|bcl|pair_pack|btp_fin)' -c 9 -o /tmp/prof_r2_layer -f $B --nelx 250 --nely 250 --warmup 0 > $O/r2_ncu_layer.log 2>&1
ncu -i /tmp/prof_r2_layer.ncu-rep --page raw --csv > $O/prof_r2_layer.raw.csv 2>/dev/null
gzip -f $O/prof_r2_n4.source.csv
du -sh $O; ls -la $O
