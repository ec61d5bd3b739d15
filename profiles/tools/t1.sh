#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_general_quads.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/t1.log
cat gpurun_out/t1.log
