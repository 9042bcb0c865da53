O=gpurun_out/${1:-r2n}; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
for k in $(seq 1 25); do
  python -m pytest tests/test_gpu_parity.py -m gpu -q -k "test_generations_bit_exact and (sab or urr or tabular)" 2>&1 | tail -30 > $O/loop_$k.log
  grep -q "failed" $O/loop_$k.log || rm $O/loop_$k.log
done
ls $O > $O/listing.txt
