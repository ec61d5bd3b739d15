#!/bin/bash
cd $GRAFT_REPO_ROOT
N=2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2c_bench_n$N.json 2> gpurun_out/r2c_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/r2c_ref_n$N.json 2> gpurun_out/r2c_ref_n$N.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c_bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value %.4g e2e %.4g stage_ms %.4f frac %.3f share %.3f clk %s'%(d['value'], d['e2e']['value'], d['roofline']['stage_ms'], d['roofline']['frac'], d['stage_only']['share_of_step'], d['clocks']['sm_mhz']), 'cross', json.dumps(d['check']['cross_n']), 'cfg5 %.4g'%d['configs'][0]['value'])
r=json.loads(open('gpurun_out/r2c_ref_n2.json').read().strip().splitlines()[-1])
print('reference arm', r.get('impl'), r.get('value'), r.get('cpu_baseline',{}).get('cores'))
P
tail -2 gpurun_out/r2c_ref_n2.err
