# Validation of the reaction-list commit on one B200: the new cases first (stop early if they fail), then the whole GPU
# parity suite, smoke() and the default bench line as the driver runs it
O=gpurun_out/${1:-r2q}; mkdir -p $O
python -m pytest tests -m gpu -q -k "reactions or tabular" 2>&1 | tail -15 > $O/reactions.log
grep -q failed $O/reactions.log && exit 1
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
( time python bench.py ) > $O/bench.json 2> $O/bench.err
