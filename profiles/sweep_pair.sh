#!/bin/bash
# sweep of the element-pair stage kernel's tuning options (run on the GPU box); prints stage ms and roofline fraction
N=${N:-500}
run() {
  echo -n "$* : "
  python bench.py --variant 0 --nelx $N --nely $N --steps 1 --warmup 1 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('stage_ms %.3f frac %.3f value %.3e power %s clk %s' % (d['roofline']['stage_ms'], d['roofline']['frac'], d['value'], d['clocks'].get('power_w_max'), d['clocks'].get('sm_mhz')))
"
}
run
run --opt pair_prefetch=0
run --opt pair_prefetch=3
run --opt pair_ne=2
run --opt pair_ne=2 --opt pair_prefetch=2
