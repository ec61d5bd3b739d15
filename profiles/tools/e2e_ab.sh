#!/bin/bash
cd $GRAFT_REPO_ROOT
for h in caller torch-pinned caller; do
python bench.py --nelx ${NEL:-1000} --nely ${NEL:-1000} --steps 2 --warmup 3 --no-cpu-baseline --no-config5 --e2e-host $h 2> gpurun_out/e2e_$h.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$h', 'value %.4g e2e %.4g e2e_ms %.1f dev_ms %.1f'%(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ms_per_step']), d['e2e']['host_buffers'])
" >> gpurun_out/e2e_ab.log
done
cat gpurun_out/e2e_ab.log
