#!/bin/bash
# 8-GPU stage time against the number of NCCL point-to-point channels (CTAs of the send/recv kernels that share the SMs with the stage kernel)
cd $GRAFT_REPO_ROOT
N=${1:-8}
: > gpurun_out/n8_nccl.log
for ch in default 2 default 1 4; do
  if [ "$ch" = "default" ]; then unset NCCL_MAX_P2P_NCHANNELS; else export NCCL_MAX_P2P_NCHANNELS=$ch; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --no-config5 --no-e2e --no-cross-check 2> gpurun_out/n8_nccl.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('channels $ch', 'value %.4g ms %.2f stage_ms %.4f share %.3f clk %s'%(d['value'], d['ms_per_step'], d['roofline']['stage_ms'], d['stage_only']['share_of_step'], d['clocks']['sm_mhz']))
" >> gpurun_out/n8_nccl.log
done
cat gpurun_out/n8_nccl.log
