#!/bin/bash
# Multi-GPU record run (under gpurun --gpus N): weak and strong scaling of the headline workload, the C4 core-sized
# HEX8 mesh and the C5 TET4 brick.   bash profiles/run_scale_r02.sh <N> <tag> [full]
N=$1; O=gpurun_out/${2:-scale}; FULL=$3; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
Q="--no-xs-leg --no-cpu-baseline"
$TR bench.py --gpus $N --steps 5 --warmup 3 $Q > $O/weak_n$N.json 2> $O/weak_n$N.err
$TR bench.py --gpus $N --steps 5 --warmup 3 $Q --no-extra --scaling strong --particles 1000000 > $O/strong1m_n$N.json 2> $O/strong1m_n$N.err
if [ -n "$FULL" ]; then
  $TR bench.py --gpus $N --steps 3 --warmup 3 $Q --no-extra --workload core_hex_15_3_17 --particles 1000000 > $O/c4_core_hex_weak_n$N.json 2> $O/c4_core_hex_weak_n$N.err
  $TR bench.py --gpus $N --steps 3 --warmup 3 $Q --no-extra --workload brick_tet_100 --particles 2000000 > $O/c5_brick_tet_100_weak_n$N.json 2> $O/c5_brick_tet_100_weak_n$N.err
  $TR bench.py --gpus $N --steps 3 --warmup 3 $Q --no-extra --workload brick_tet_100 --particles 2000000 --scaling strong > $O/c5_brick_tet_100_strong_n$N.json 2> $O/c5_brick_tet_100_strong_n$N.err
fi
python -m pytest tests -m gpu -x -q -k "two_gpus" > $O/pytest_n$N.log 2>&1
