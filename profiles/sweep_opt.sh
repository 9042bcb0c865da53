#!/bin/bash
# sweep one bench.py option:  profiles/sweep_opt.sh out.txt "<fixed flags>" --tail-threshold 0 65536 ...
out=$1; fixed=$2; opt=$3; shift 3
for v in "$@"; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-xs-leg --no-extra $fixed $opt $v 2>/dev/null \
    | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); r=d['roofline']; print('$fixed $opt $v |', round(d['value']), 'hist/s', round(d['ms_per_step'],2), 'ms/step track', round(r['kernel_share_of_step']*d['ms_per_step'],2), 'coll', round(r['other_kernels_ms_per_step']['collide+fallback'],2), 'rounds', d['stats']['event_rounds_per_batch'], 'launches', d['gpu_launches']/d['steps'], 'k', d['stats']['keff'], 'e2e', round(d['e2e']['value']))" >> $out
done
cat $out
