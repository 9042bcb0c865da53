O=gpurun_out/r2l; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
bash profiles/sweep_opt.sh $O/bench_quick.txt "" --sync-every 4 > /dev/null 2>&1
bash profiles/sweep_opt.sh $O/bench_quick.txt "--particles 200000" --sync-every 4 > /dev/null 2>&1
bash profiles/sweep_opt.sh $O/bench_quick.txt "--workload pin_cell_tet" --sync-every 4 > /dev/null 2>&1
bash profiles/sweep_opt.sh $O/bench_quick.txt "--opt force_rare=1" --sync-every 4 > /dev/null 2>&1
bash profiles/sweep_opt.sh $O/bench_quick.txt "--workload pin_cell_tet --opt force_rare=1" --sync-every 4 > /dev/null 2>&1
