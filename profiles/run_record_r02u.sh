#!/bin/bash
# Round-2 closing record on one B200 (under gpurun): the newest parity cases first (stop early if they fail), the whole GPU
# parity suite, smoke(), the default bench line as the driver runs it, the reference arm, and an ncu launch list.
O=gpurun_out/${1:-r2u}; mkdir -p $O
python -m pytest tests -m gpu -q -k "evaporation or fission" 2>&1 | tail -25 > $O/new_cases.log
grep -q failed $O/new_cases.log && exit 1
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
( time python bench.py ) > $O/bench.json 2> $O/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/ref.json 2> $O/ref.err
Q="--no-extra --no-xs-leg --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 $Q > $O/launches.log 2>&1
