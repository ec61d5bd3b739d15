#!/bin/bash
# Round-2 final evidence (one B200): full GPU test suite, driver-style bench line, launch list at 250x250, ncu --set full of the stage
# kernel at nop 4 (500x500) and nop 8 (200x200, 10 layers), CSV exports only (the reports stay on the box).
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2final; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/gpu_tests.log
python bench.py --steps 3 --warmup 3 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
NCU="ncu --clock-control none"
B="python bench.py --steps 1 --no-cpu-baseline --no-e2e --no-config5 --no-cross-check"
$NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file $O/r2_launches_250x250.csv $B --nelx 250 --nely 250 --warmup 1 > $O/ncu_l.log 2>&1
$NCU --set full --import-source on -k regex:k_btp_stage_pair -s 30 -c 1 -o /tmp/prof_r2_n4 -f $B --nelx 500 --nely 500 --warmup 0 > $O/ncu_n4.log 2>&1
ncu -i /tmp/prof_r2_n4.ncu-rep --page raw --csv > $O/prof_r2_n4.raw.csv 2>/dev/null
ncu -i /tmp/prof_r2_n4.ncu-rep --page source --csv --print-source sass > $O/prof_r2_n4.sass.csv 2>/dev/null
gzip -f $O/prof_r2_n4.sass.csv
$NCU --set full -k regex:k_btp_stage_pair -s 30 -c 1 -o /tmp/prof_r2_n8 -f $B --nop 8 --layers 10 --nelx 200 --nely 200 --warmup 0 > $O/ncu_n8.log 2>&1
ncu -i /tmp/prof_r2_n8.ncu-rep --page raw --csv > $O/prof_r2_n8.raw.csv 2>/dev/null
cat $O/gpu_tests.log; tail -c 600 $O/r2_bench_n1.json; ls -la $O
