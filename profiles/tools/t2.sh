#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "layer_phase or step_parity or partitioned_equals" 2>&1 | tail -3 > gpurun_out/t2.log
bash profiles/tools/layer_ab.sh libhnumo_b200.so >> gpurun_out/t2.log 2>&1
for o in 1 0 1 0; do python bench.py --nelx 500 --nely 500 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-config5 --no-cross-check --opt mom_volume_batched=$o 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('batched=$o ms_per_step %.2f share %.4f'%(d['ms_per_step'], d['stage_only']['share_of_step']))
" >> gpurun_out/t2.log; done
cat gpurun_out/t2.log
