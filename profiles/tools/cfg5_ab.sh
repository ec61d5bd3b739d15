#!/bin/bash
# config 5 (scaled: 250x250 elements, nop 8, 10 layers), whole step, for several library builds
cd $GRAFT_REPO_ROOT
: > gpurun_out/cfg5_ab.log
for r in 1 2; do for lib in "$@"; do
  HNUMO_LIB_PATH=$PWD/h-numo_b200/$lib python bench.py --nop 8 --layers 10 --nelx 250 --nely 250 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-config5 --no-cross-check 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib', 'value %.4g ms %.1f stage_ms %.4f share %.3f'%(d['value'], d['ms_per_step'], d['roofline']['stage_ms'], d['stage_only']['share_of_step']))
" >> gpurun_out/cfg5_ab.log
done; done
cat gpurun_out/cfg5_ab.log
