#!/bin/bash
# sweep one bench.py option:  profiles/sweep_opt.sh out.txt --tail-threshold 0 65536 ...
out=$1; opt=$2; shift 2
for v in "$@"; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline $opt $v 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$opt $v', round(d['value']), 'hist/s', round(d['ms_per_step'],2), 'ms/step track', round(r['kernel_share_of_step']*d['ms_per_step'],2), 'other', r['other_kernels_ms_per_step'], 'rounds', d['config']['event_rounds_per_batch'], 'k', d['config']['keff'], 'e2e', round(d['e2e']['value']))" >> $out
done
cat $out
