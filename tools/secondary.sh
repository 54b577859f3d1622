#!/bin/bash
# secondary workloads of BASELINE.json (not the headline): k=5 generic vs brick path, Kershaw smoother step, weak-scaling size per GPU
run() { echo "== $*"; timeout 600 python bench.py "$@" --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d.get('extra',{})
print(d['config']['workload'][38:110], d['dtype'], 'n=%.3g'%d['config']['n_dofs'], 'step %.3e DoF/s'%d['value'], ' '.join('%s=%.3e'%(k,v) for k,v in e.items()))"; }
run --degree 5 --cells 92,92,92
DASM_FORCE_GENERIC=1 run --degree 5 --cells 92,92,92
run --degree 4 --cells 160,160,160
run --degree 4 --cells 96,96,96 --map kershaw --mapping-type "quadratic geometry"
run --degree 4 --cells 96,96,96 --map kershaw --mapping-type "merged"
run --degree 4 --cells 96,96,96 --map kershaw
