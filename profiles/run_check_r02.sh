O=gpurun_out/${1:-r2o}; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
for k in 1 2 3 4 5 6; do
  python -m pytest tests/test_gpu_parity.py -m gpu -q -k "test_generations_bit_exact and (sab or urr or full)" 2>&1 | tail -3 > $O/loop_$k.log
done
Q="--no-extra --no-xs-leg --no-cpu-baseline"
bash profiles/sweep_opt.sh $O/bench_quick.txt "" --sync-every 4 > /dev/null 2>&1
bash profiles/sweep_opt.sh $O/bench_quick.txt "--particles 200000" --sync-every 4 > /dev/null 2>&1
bash profiles/sweep_opt.sh $O/bench_quick.txt "--workload pin_cell_tet" --sync-every 4 > /dev/null 2>&1
bash profiles/sweep_opt.sh $O/bench_quick.txt "--tally TRACKLENGTH" --sync-every 4 > /dev/null 2>&1
python bench.py --workload core_hex_15_3_17 --particles 1000000 --steps 3 --warmup 3 $Q > $O/c4_core_hex_n1.json 2> $O/c4_core_hex_n1.err
