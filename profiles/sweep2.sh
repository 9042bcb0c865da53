#!/bin/bash
# profiles/sweep2.sh out.txt "<fixed opts>" --optA "a1 a2" --optB "b1 b2": full grid of two bench.py options
out=$1; fixed=$2; oa=$3; va=$4; ob=$5; vb=$6
for a in $va; do for b in $vb; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline $fixed $oa $a $ob $b 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$oa $a $ob $b', round(d['value']), 'hist/s', round(d['ms_per_step'],2), 'ms/step track', round(r['kernel_share_of_step']*d['ms_per_step'],2), 'coll', round(r['other_kernels_ms_per_step']['collide+fallback'],2), 'rounds', d['config']['event_rounds_per_batch'], 'k', d['config']['keff'])" >> $out
done; done
cat $out
