#!/bin/bash
# A/B of library build variants (macaw_b200/build.py build_variant): same bench, different library.
#   profiles/ab_variants.sh out.txt "<bench flags>" name1 name2 ...   ("default" = the in-tree library)
out=$1; fixed=$2; shift 2
for v in "$@"; do
  lib=$PWD/macaw_b200/variants/libmacaw_b200_$v.so; [ "$v" = default ] && lib=$PWD/macaw_b200/libmacaw_b200.so
  MCW_LIB_PATH=$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-xs-leg --no-extra $fixed 2>/dev/null \
    | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); r=d['roofline']; print('$v $fixed |', round(d['value']), 'hist/s', round(d['ms_per_step'],2), 'ms/step track', round(r['kernel_share_of_step']*d['ms_per_step'],2), 'coll', round(r['other_kernels_ms_per_step']['collide+fallback'],2), 'rounds', d['stats']['event_rounds_per_batch'], 'k', d['stats']['keff'], d['stats']['library_loaded'])" >> $out
done
cat $out
