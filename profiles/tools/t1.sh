#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest ${T1:-tests/test_gpu_edge_cases.py} -m gpu -q 2>&1 | tail -40 > gpurun_out/t1.log
cat gpurun_out/t1.log
