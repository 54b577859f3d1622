for d in 0 1 2 4 6 8 16 30 31; do
  echo "== DBG $d"
  DASM_FAST_DBG=$d DASM_FAST_PROF=1 timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2> /tmp/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernel_time_ms'], d['clocks'])"
  grep "fast prof" /tmp/err.txt | tail -2 | cut -c1-330
done
