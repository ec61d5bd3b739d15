#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2final; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/gpu_tests.log
python bench.py --steps 3 --warmup 3 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
cat $O/gpu_tests.log $O/smoke.log; tail -c 400 $O/r2_bench_n1.json
