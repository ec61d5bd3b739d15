#!/bin/bash
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_step_parity or test_phase_parity" 2>&1 | tail -3 > gpurun_out/exp1_tests.log
B="python bench.py --nelx 500 --nely 500 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-config5 --no-cross-check"
for pt in 128 256 64; do
  HNUMO_PACK_THREADS=$pt $B > gpurun_out/exp1_b_$pt.json 2> gpurun_out/exp1_b_$pt.err
done
HNUMO_PACK_THREADS=256 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/exp1_launches_256.csv python bench.py --nelx 250 --nely 250 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-config5 --no-cross-check > gpurun_out/exp1_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/exp1_launches_128.csv python bench.py --nelx 250 --nely 250 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-config5 --no-cross-check > gpurun_out/exp1_ncu2.log 2>&1
cat gpurun_out/exp1_tests.log
