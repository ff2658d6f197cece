for cb in 0 1; do echo COL_BLOCKS=$cb; FDR_COL_BLOCKS=$cb python bench.py --workload rgb16384 --steps 3 --warmup 2 --no-cpu-baseline --no-check --no-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], {k:round(v['ms_per_step_summed'],3) for k,v in d['roofline']['in_step']['kernels'].items()})
"; done
python bench.py --workload rgb4096 --steps 10 --warmup 3 --no-cpu-baseline --no-check --no-e2e --flush-l2 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('rgb4096', d['value'], d['ms_per_step'], {k:round(v['ms_per_step_summed'],3) for k,v in d['roofline']['in_step']['kernels'].items()})
"
