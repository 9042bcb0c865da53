# Validation of the temperature-set commit on one B200: the new cases first (stop early if they fail), then the whole GPU
# parity suite, smoke() and the default bench line as the driver runs it
O=gpurun_out/${1:-r2s}; mkdir -p $O
python -m pytest tests -m gpu -q -k "temperatures" 2>&1 | tail -25 > $O/temperatures.log
grep -q failed $O/temperatures.log && exit 1
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
( time python bench.py ) > $O/bench.json 2> $O/bench.err
