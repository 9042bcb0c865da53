"""Multi-rank plumbing of the particle-sharded eigenvalue iteration, on torch.distributed
(NCCL over NVLink on the GPU box, gloo in the CPU tests).

Work shards by particle exactly like OpenMC's native MPI mode (openmc/src/simulation.cpp:529-551
calculate_work): rank r owns source-bank indices [work_index[r], work_index[r+1]); mesh and nuclear
data are replicated.  Per generation there is one exchange step, the restatement of
openmc/src/eigenvalue.cpp:71-311 synchronize_bank with its MPI calls mapped as

    MPI_Exscan + MPI_Bcast of bank sizes (eigenvalue.cpp:84-96, MaCaw OpenMCStudy.C:467-477)
        -> one all_gather of int64 counts, prefix sum on every rank
    MPI_Exscan + MPI_Allgather of sampled counts (eigenvalue.cpp:170-176)
        -> one all_gather of int64 counts
    MPI_Isend / MPI_Irecv of SourceSite ranges (eigenvalue.cpp:212-300)
        -> batched isend/irecv of 96-byte site ranges between device buffers
    MPI_Allreduce of k_generation and MPI_Reduce of tallies (eigenvalue.cpp:60, tally.cpp:761-798)
        -> one all_reduce(SUM, float64) over [tally VALUEs | k_col k_abs k_trk leak | weight]

The functions below only see tensors and integer plans, so the same code runs on CPU tensors in the
world_size-2 gloo tests."""
from __future__ import annotations

import numpy as np

SITE_BYTES = 96


def calculate_work(n_particles: int, nranks: int) -> np.ndarray:
    """OpenMC calculate_work (simulation.cpp:529-551) -- pure host arithmetic."""
    min_work, rem = divmod(int(n_particles), int(nranks))
    wi = np.zeros(nranks + 1, np.int64)
    for i in range(nranks):
        wi[i + 1] = wi[i] + (min_work + 1 if i < rem else min_work)
    return wi


def exclusive_offsets(counts) -> np.ndarray:
    c = np.asarray(counts, np.int64)
    out = np.zeros(c.size + 1, np.int64)
    np.cumsum(c, out=out[1:])
    return out


def exchange_plan(sampled_counts, work_index):
    """Who sends which range to whom.  Rank s holds global sampled positions
    [off[s], off[s+1]); rank d must end up with [work_index[d], work_index[d+1]).
    Returns a list of (src, dst, src_local_offset, dst_local_offset, count)."""
    off = exclusive_offsets(sampled_counts)
    wi = np.asarray(work_index, np.int64)
    plan = []
    for s in range(len(sampled_counts)):
        for d in range(len(wi) - 1):
            lo, hi = max(off[s], wi[d]), min(off[s + 1], wi[d + 1])
            if hi > lo:
                plan.append((s, d, int(lo - off[s]), int(lo - wi[d]), int(hi - lo)))
    return plan


def all_gather_int(value: int, group=None, device="cpu"):
    """One int64 per rank -> list on every rank: one collective into one tensor and ONE device -> host copy
    (the per-rank `.item()` reads of a tensor list cost a stream synchronisation each)."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    out = torch.empty(ws, dtype=torch.int64, device=device)
    try:
        dist.all_gather_into_tensor(out, t, group=group)
    except (RuntimeError, NotImplementedError):  # a backend without the flat variant
        parts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(ws)]
        dist.all_gather(parts, t, group=group)
        out = torch.cat(parts)
    return [int(x) for x in out.cpu().tolist()]


def warm_up(group=None, device="cpu"):
    """Open every point-to-point channel and the collective rings once, outside any timed region: NCCL
    connects a pair of ranks lazily on their first send / recv, and which pairs exchange sites depends
    on the generation (at 8 ranks new pairs kept appearing for dozens of generations, ~100 ms each)."""
    import torch
    import torch.distributed as dist
    rank, ws = dist.get_rank(group), dist.get_world_size(group)
    if ws == 1:
        return
    send = torch.zeros(ws, 96, dtype=torch.uint8, device=device)
    recv = torch.zeros(ws, 96, dtype=torch.uint8, device=device)
    ops = []
    for peer in range(ws):
        if peer != rank:
            ops.append(dist.P2POp(dist.isend, send[peer], peer, group))
            ops.append(dist.P2POp(dist.irecv, recv[peer], peer, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    t = torch.zeros(8, dtype=torch.float64, device=device)
    dist.all_reduce(t, group=group)
    all_gather_int(0, group, device)


def exchange_sites(sampled_u8, source_u8, sampled_counts, work_index, group=None):
    """Move sampled sites to their owners.  sampled_u8 / source_u8 are flat uint8 tensors over this
    rank's sampled-site and source-bank buffers (device or CPU)."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    ops = []
    for s, d, so, do, n in exchange_plan(sampled_counts, work_index):
        if s == rank and d == rank:
            source_u8[do * SITE_BYTES:(do + n) * SITE_BYTES].copy_(sampled_u8[so * SITE_BYTES:(so + n) * SITE_BYTES])
        elif s == rank:
            ops.append(dist.P2POp(dist.isend, sampled_u8[so * SITE_BYTES:(so + n) * SITE_BYTES], d, group))
        elif d == rank:
            ops.append(dist.P2POp(dist.irecv, source_u8[do * SITE_BYTES:(do + n) * SITE_BYTES], s, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def synchronize_banks(engine, batch, rank, nranks, work_index, group=None, device="cpu"):
    """postExecuteStudy() for nranks > 1 (MaCaw OpenMCStudy.C:404-607, OpenMC eigenvalue.cpp:71-311):
    bank offsets, the fused tally / k all-reduce, bank sampling on every rank's slice of the global bank,
    last-rank trim / pad, redistribution of the sampled sites to their owners, finalize.

    `engine` is the per-rank compute object (the GPU Context through study.DeviceEngine, or the CPU
    oracle in the gloo tests).  It must offer bank_count(), set_bank_window(start, total),
    reduction_tensor() -> float64 tensor reduced in place, sample_bank(batch) -> n,
    finish_sampling(start, is_last) -> n, sampled_tensor(n) / source_tensor() -> uint8 tensors,
    set_source_count(n), finalize_generation(batch).

    On a GPU the caller runs this with the library's stream as torch's current stream.  The all-reduce is
    issued first and asynchronously: NCCL runs it beside the bank-sampling kernels and the site exchange, and
    the stream only waits for it before finalize_generation reads the reduced k sums -- nothing here blocks
    the host except the two count all-gathers, which the exchange plan needs as host integers."""
    import torch.distributed as td
    red = engine.reduction_tensor()
    red_work = td.all_reduce(red, op=td.ReduceOp.SUM, group=group, async_op=True)
    counts = all_gather_int(engine.bank_count(), group, device)
    if min(counts) == 0:
        red_work.wait()
        raise RuntimeError(f"No fission sites banked on rank {counts.index(0)}")  # eigenvalue.cpp:107-109
    off = exclusive_offsets(counts)
    engine.set_bank_window(int(off[rank]), int(off[-1]))
    n_s = engine.sample_bank(batch)
    scounts = all_gather_int(n_s, group, device)
    soff = exclusive_offsets(scounts)
    n_s = engine.finish_sampling(int(soff[rank]), rank == nranks - 1)
    # only the last rank trims / pads (eigenvalue.cpp:184-206) and it always ends at exactly n_particles,
    # so every rank can derive the final counts without a second all-gather
    scounts[-1] = max(0, int(work_index[-1]) - int(soff[nranks - 1]))
    assert rank != nranks - 1 or n_s == scounts[-1], (n_s, scounts)
    exchange_sites(engine.sampled_tensor(max(scounts[rank], 0)), engine.source_tensor(), scounts, work_index, group)
    engine.set_source_count(int(work_index[rank + 1] - work_index[rank]))
    red_work.wait()  # a stream-level wait on a GPU
    return engine.finalize_generation(batch)


class DeviceBuffer:
    """Zero-copy torch view of a raw device pointer handed out by the C ABI."""

    def __init__(self, ptr, nbytes, typestr="|u1", itemsize=1):
        self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def device_tensor(ptr, count, kind="u8", device=None):
    import torch
    if kind == "u8":
        buf = DeviceBuffer(ptr, max(count, 1))
    else:
        buf = DeviceBuffer(ptr, max(count, 1) * 8, "<f8", 8)
    return torch.as_tensor(buf, device=device if device is not None else "cuda")
