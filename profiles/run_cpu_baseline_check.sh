# The CPU port's speed on the GPU box's host (the quiet machine the cpu_baseline is quoted on): reference arm, 2 steps
O=gpurun_out/${1:-r2v}; mkdir -p $O
python bench.py --impl reference --steps 2 --warmup 1 > $O/ref.json 2> $O/ref.err
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generations_bit_exact and (pin_cell_tet- or beavrs_assembly- or temperatures or fission)" > $O/pytest_subset.log 2>&1
