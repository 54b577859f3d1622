#!/bin/bash
# degree / number-type sweep of the smoother step (BASELINE.json configs[1]): ~1e8 DoFs per run, one JSON line each
out=${1:-gpurun_out/sweep.jsonl}
: > $out
for number in double float; do
  for k in 1 2 3 4 5 6; do
    case $k in
      1) cells=464,464,464;; 2) cells=232,232,232;; 3) cells=156,156,156;; 4) cells=192,128,64;; 5) cells=92,92,92;; 6) cells=76,76,76;;
    esac
    timeout 300 python bench.py --degree $k --number $number --cells $cells --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --extra 2>/dev/null >> $out || echo "{\"degree\": $k, \"number\": \"$number\", \"failed\": true}" >> $out
  done
done
python - <<'PY' $out
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l)
    if d.get('failed'): print(d); continue
    e=d.get('extra',{})
    print(d['config']['workload'][38:60], d['dtype'], 'n=%.3g'%d['config']['n_dofs'], 'step %.3e DoF/s'%d['value'], 'frac %.3f'%d['roofline']['step_frac_of_hbm_roofline'], ' '.join('%s=%.3e'%(k,v) for k,v in e.items()))
PY
