#!/bin/bash
# Round-2 record run on one B200 (under gpurun): GPU parity tests, the default bench line, the reference arm,
# a launch list, ncu --set full captures of the tracking kernel (collision / track-length tally) and the lookup leg,
# exported to CSV on the box (the .ncu-rep files stay there: gpurun_out/ is limited to 64 MiB).
O=gpurun_out/${1:-r2k}
mkdir -p $O
export_rep() { # <name>: raw + source pages as CSV, then drop the report
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  ncu -i $O/$1.ncu-rep --page source --print-source cuda,sass --csv > $O/$1_src.csv 2>/dev/null || \
    ncu -i $O/$1.ncu-rep --page source --csv > $O/$1_src.csv 2>/dev/null
  rm -f $O/$1.ncu-rep
}
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
( time python bench.py ) > $O/bench.json 2> $O/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/ref.json 2> $O/ref.err
python bench.py --tally TRACKLENGTH --no-extra --no-xs-leg --no-cpu-baseline --steps 5 --warmup 3 > $O/bench_tracklength.json 2> $O/bench_tracklength.err
Q="--no-extra --no-xs-leg --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 $Q > $O/launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_track -s 3 -c 1 -o $O/track python bench.py --steps 1 --warmup 0 $Q > $O/ncu_track.log 2>&1
export_rep track
RED=l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_requests_op_red.sum,l1tex__t_requests_pipe_lsu_mem_global_op_red.sum,smsp__inst_executed_op_global_red.sum
ncu --set full --metrics $RED --clock-control none --import-source on -k regex:k_track -s 3 -c 1 -o $O/track_tl python bench.py --tally TRACKLENGTH --steps 1 --warmup 0 $Q > $O/ncu_track_tl.log 2>&1
export_rep track_tl
ncu --set full --clock-control none -k regex:k_xs_lookup -c 6 -o $O/xs python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline > $O/ncu_xs.log 2>&1
export_rep xs
# A/B of build variants: fused-tail L1 prefetch (pf1, pf2), out-of-line collision math (ool), both (oolpf2)
for fl in "" "--particles 200000" "--workload pin_cell_tet"; do
  bash profiles/ab_variants.sh $O/ab_variants.txt "$fl" default pf1 pf2 ool oolpf2 default > /dev/null 2>&1
done
# meshes larger than L2: lexicographic vs Morton element numbering
for w in brick_tet_100 brick_tet_100_morton brick_tet_160 brick_tet_160_morton; do
  python bench.py --workload $w --particles 2000000 --steps 3 --warmup 3 $Q > $O/bench_$w.json 2> $O/bench_$w.err
done
du -sh $O > $O/size.txt
