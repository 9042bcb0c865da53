"""World-size-N NCCL worker for tests/test_gpu_parity.py::test_two_gpus_equal_one_gpu: one rank = one B200 =
one mcw_handle, driven through macaw_b200.study.OpenMCStudy exactly as bench.py drives it; writes what rank r
ended up with to <out>.<r>.npz.

  python tests/dist_gpu_worker.py <rank> <world> <port> <out> <problem> <n_particles> <n_batches>
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, port, out, name, n_particles, n_batches = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]),
                                                           sys.argv[4], sys.argv[5], int(sys.argv[6]), int(sys.argv[7]))
    import torch
    import torch.distributed as td
    from macaw_b200 import cstructs, study as st
    from tests import problems
    torch.cuda.set_device(rank)
    td.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                          device_id=torch.device("cuda", rank))
    mesh, lib, bm, s = problems.PROBLEMS[name](n_particles=n_particles)
    blocks = list(range(len(bm))) if int(mesh.block_id.max()) < len(bm) else [int(b) for b in mesh.block_ids_sorted()]
    kernel = st.CollisionKernel(blocks=blocks, materials=[m + 1 for m in bm], z_coord=100.0 if mesh.dim == 2 else None)
    tally = st.OpenMCTally(id=1, scores=["flux", "nu-fission"], filters=["cell"], estimator="COLLISION")
    study = st.OpenMCStudy(mesh, lib, kernel, s, [tally], device=rank, rank=rank, nranks=world)
    ks = []
    for _ in range(n_batches):
        r = study.execute()
        ks.append([r.k_generation, r.keff, r.k_collision, r.k_absorption, r.k_tracklength])
    t_sum, _, _ = study.ctx.get_tally(1)
    np.savez(f"{out}.{rank}.npz", source=study.ctx.source_bank(), ks=np.asarray(ks), tally_sum=t_sum,
             records=study.ctx.history_records())
    td.destroy_process_group()


if __name__ == "__main__":
    main()
