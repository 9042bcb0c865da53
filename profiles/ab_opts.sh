#!/bin/bash
# profiles/ab_opts.sh out.txt "<workload args>" "optset1" "optset2" ...   (each optset: bench.py flags)
out=$1; fixed=$2; shift 2
for o in "$@"; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-xs-leg $fixed $o 2>/dev/null \
    | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); r=d['roofline']; print('$fixed | $o |', round(d['value']), 'hist/s', round(d['ms_per_step'],2), 'ms/step track', round(r['kernel_share_of_step']*d['ms_per_step'],2), 'coll', round(r['other_kernels_ms_per_step']['collide+fallback'],2), 'rounds', d['config']['event_rounds_per_batch'], 'k', d['config']['keff'])" >> $out
done
cat $out
