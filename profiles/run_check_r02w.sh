# The GPU lookup against the reference's outputs at several temperatures (tests/golden/ref_collisions_full.npz), and the
# temperature cases again after mcw_set_temperature learned to prepare again
O=gpurun_out/${1:-r2w}; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "reference_fixture or reference_objects_fixture or temperatures or full_fission_bank" 2>&1 | tail -15 > $O/pytest_subset.log
