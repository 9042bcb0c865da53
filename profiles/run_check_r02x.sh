# S(a,b) tables at several temperatures under a temperature field: GPU parity cases and the C++ host
O=gpurun_out/${1:-r2x}; mkdir -p $O
python -m pytest tests -m gpu -q -k "sab_temperatures or (generations_bit_exact and (pin_cell_tet_sab or temperatures))" 2>&1 | tail -15 > $O/pytest_subset.log
