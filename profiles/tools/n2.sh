#!/bin/bash
cd $GRAFT_REPO_ROOT
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --no-config5 > gpurun_out/r2b_bench_n$N.json 2> gpurun_out/r2b_bench_n$N.err
python - <<'P'
import json,sys,glob
for f in sorted(glob.glob('gpurun_out/r2b_bench_n*.json')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,e); continue
    print(f, 'value %.4g e2e %.4g ms %.1f e2e_ms %.1f frac %.3f share %.3f numa %s cross %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['stage_only']['share_of_step'], d['run'].get('host_numa_node_rank0'), json.dumps(d['check'].get('cross_n',{}).get('inviscid'))))
P
tail -3 gpurun_out/r2b_bench_n$N.err
