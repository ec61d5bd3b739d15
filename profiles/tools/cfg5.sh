#!/bin/bash
# BASELINE config 5 as the main leg (nop 8, 10 layers, 500x500 elements), with the end-to-end leg
cd $GRAFT_REPO_ROOT
python bench.py --nop 8 --layers 10 --nelx 500 --nely 500 --steps 2 --warmup 2 --no-cpu-baseline --no-config5 > gpurun_out/r2_bench_config5_n1.json 2> gpurun_out/r2_bench_config5_n1.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_config5_n1.json').read().strip().splitlines()[-1])
print('config5 value %.4g e2e %.4g ms %.1f stage_ms %.4f frac %.3f traffic %.4g share %.3f clk %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['stage_ms'], d['roofline']['frac'], d['roofline']['traffic'] or 0, d['stage_only']['share_of_step'], d['clocks']['sm_mhz']))
P
tail -3 gpurun_out/r2_bench_config5_n1.err
