#!/bin/bash
# A/B of k_track build variants (macaw_b200/build.py build_variant): same bench, different library.
#   profiles/ab_variants.sh out.txt name1 name2 ...
out=$1; shift
for v in "$@"; do
  MCW_LIB_PATH=$PWD/macaw_b200/variants/libmacaw_b200_$v.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$v', round(d['value']), 'hist/s', round(d['ms_per_step'],2), 'ms/step track', round(r['kernel_share_of_step']*d['ms_per_step'],2), 'other', r['other_kernels_ms_per_step'], 'k', d['config']['keff'])" >> $out
done
cat $out
