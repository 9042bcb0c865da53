"""World-size-N gloo worker for tests/test_host_cpu.py: drives one CPU-oracle problem per rank through
macaw_b200.dist.synchronize_banks (the same host code the GPU study runs over NCCL) and writes what
rank r ended up with to <out>.<r>.npz.

  python tests/dist_worker.py <rank> <world> <port> <out> <n_particles> <n_batches> [problem]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class OracleEngine:
    """dist.synchronize_banks hooks over the CPU oracle (host tensors)."""

    def __init__(self, o, n_local_max):
        import ctypes as C
        self.o, self.C = o, C
        self.source = np.zeros(n_local_max * 96, np.uint8)
        self._sampled = None

    def bank_count(self):
        return int(self.o.fission_bank().size)

    def set_bank_window(self, start, total):
        self.o.L.orc_set_bank_window(self.o.h, start, total)

    def reduction_tensor(self):
        import torch
        rec = self.o.history_records()
        w = self.o.source_bank()["wgt"].sum()
        v1, _, _, _ = self.o.get_tally(1)
        self.red = torch.tensor(np.concatenate([v1, [rec["k_collision"].sum(), rec["k_absorption"].sum(),
                                                     rec["k_tracklength"].sum(), 0.0, w]]), dtype=torch.float64)
        return self.red

    def sample_bank(self, batch):
        return int(self.o.L.orc_sample_bank(self.o.h, batch))

    def finish_sampling(self, start, is_last):
        return int(self.o.L.orc_finish_sampling(self.o.h, start, int(is_last)))

    def sampled_tensor(self, n):
        import torch
        s = self.o.sampled()
        self._sampled = np.frombuffer(s.tobytes() + b"\0", np.uint8).copy()
        return torch.from_numpy(self._sampled)

    def source_tensor(self):
        import torch
        return torch.from_numpy(self.source)

    def set_source_count(self, n):
        from macaw_b200.cstructs import SITE_DTYPE
        self.o.set_source_bank(self.source[: n * 96].view(SITE_DTYPE))

    def finalize_generation(self, batch):
        from macaw_b200.cstructs import GenerationResult
        res = GenerationResult()
        g = self.red[-5:].numpy().copy()
        self.o.L.orc_finalize_generation(self.o.h, batch, g.ctypes.data, self.C.byref(res))
        return res


def main():
    rank, world, port, out, n_particles, n_batches = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]),
                                                     sys.argv[4], int(sys.argv[5]), int(sys.argv[6]))
    import torch.distributed as td
    from macaw_b200 import dist
    from oracle import bridge
    from tests import problems
    td.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    dist.warm_up(None, "cpu")  # the channel warm-up the GPU study does at construction
    name = sys.argv[7] if len(sys.argv) > 7 else "pin_cell_tet"
    mesh, lib, bm, s = problems.PROBLEMS[name](n_particles=n_particles)
    o = problems.configure(bridge.Oracle(rank=rank, nranks=world), mesh, lib, bm, s, True, False)
    wi = dist.calculate_work(n_particles, world)
    o.sample_initial_source()
    eng = OracleEngine(o, int((wi[1:] - wi[:-1]).max()))
    ks = []
    for b in range(1, n_batches + 1):
        o.transport_generation(b)
        res = dist.synchronize_banks(eng, b, rank, world, wi, None, "cpu")
        ks.append([res.k_generation, res.keff, res.k_collision, res.k_absorption, res.k_tracklength])
    _, t_sum, _, _ = o.get_tally(1)
    np.savez(f"{out}.{rank}.npz", source=o.source_bank(), ks=np.asarray(ks), tally_sum=t_sum, wi=wi)
    td.destroy_process_group()


if __name__ == "__main__":
    main()
