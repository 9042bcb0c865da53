# Validation of the fission-data commit on one B200: the new cases first (stop early if they fail), then the whole GPU
# parity suite, smoke() and the default bench line as the driver runs it
O=gpurun_out/${1:-r2t}; mkdir -p $O
python -m pytest tests -m gpu -q -k "fission or temperatures" 2>&1 | tail -25 > $O/new_cases.log
grep -q failed $O/new_cases.log && exit 1
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
( time python bench.py ) > $O/bench.json 2> $O/bench.err
