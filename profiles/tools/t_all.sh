#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/t_all.log 2>&1
cat gpurun_out/t_all.log
