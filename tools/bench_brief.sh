#!/bin/bash
# usage: tools/bench_brief.sh [lib] -- prints ms/step, frames/s, fwd ms, bwd ms
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.4f  frames/s %.3e  fwd %.4f  bwd %.4f' % (d['ms_per_step'], d['value'], d['roofline']['fwd']['ms'], d['roofline']['bwd']['ms']))
    elif l.strip(): print(l[:160])
"
