#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/exp4_tests.log
cat gpurun_out/exp4_tests.log
bash profiles/tools/ab.sh 500 3 libhnumo_b200_base.so libhnumo_b200.so
