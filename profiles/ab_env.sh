#!/bin/bash
# A/B of an environment switch on one workload: profiles/ab_env.sh VAR "v1 v2 ..." [bench.py args...]
# prints ms/step, per-pass ms and GB/s, parity counts for every value ("-" = variable unset).
var=$1; shift
vals=$1; shift
for v in $vals; do
  if [ "$v" = "-" ]; then unset $var; else export $var=$v; fi
  timeout 300 python bench.py "$@" --no-cpu-baseline --no-side --no-e2e --no-check 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline'].get('in_step',{}).get('kernels',{})
print('$var=$v', 'ms/step %.4f' % d['ms_per_step'], {n:(round(x['ms_per_step_summed'],4),round(x['GBps_per_launch'])) for n,x in k.items()}, 'launches', d['gpu_launches']//d['steps'], d['parity'] and (d['parity']['off_by_1'], d['parity']['off_by_more']))"
done
