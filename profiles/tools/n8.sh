#!/bin/bash
# driver-style scaling leg: bench.py on N GPUs of one box over NCCL (usage: n8.sh N)
cd $GRAFT_REPO_ROOT
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 3000 gpurun_out/r2_bench_n$N.json; tail -5 gpurun_out/r2_bench_n$N.err
