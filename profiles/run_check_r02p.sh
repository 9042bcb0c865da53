# Validation of the secondary-law commit on one B200: GPU parity suite, then the default bench line as the driver runs it
O=gpurun_out/${1:-r2p}; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tabular" 2>&1 | tail -5 > $O/tabular.log
( time python bench.py ) > $O/bench.json 2> $O/bench.err
