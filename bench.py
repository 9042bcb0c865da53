#!/usr/bin/env python
"""bench.py -- neutron histories/sec over active batches (BASELINE.json metric).

A "step" is one batch = one generation of the k-eigenvalue iteration.  Workload of the headline line:
BASELINE.json configs[2], the BEAVRS 17x17 assembly on the reference's own mesh (31enr_16_in.e, 23 876
QUAD4, 35 synthetic nuclides), 1 M histories per GPU per batch (weak scaling; `--scaling strong` fixes the
global count instead).  The same JSON line also carries

  reference_size : the same workload at the reference's own size (examples/beavrs_assembly/settings.xml:4-6,
                   200 000 particles x 5 batches, 0 inactive), a fresh study, all five batches timed
  extra          : BASELINE.json configs[1] (UO2 pin cell, 101 400 TET4, 9 nuclides), N = 1 only
  xs_lookup      : the stand-alone cross-section lookup (BASELINE.json's second metric), N = 1 only

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload W] [--particles P]

value = whole-job histories/s with the source bank resident in HBM; e2e = the same through the C ABI with
HOST (pinned) source banks uploaded and the next-generation bank + k read back every step; roofline = the
tracking kernel's algorithmic bytes over its CUDA-event time inside the timed region, with the
ncu-measured DRAM traffic and FP64-pipe share beside it (profiles/); cpu_baseline = the CPU oracle
(history-based restatement of the reference algorithms) on this box's host cores, core count stated.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL writes its version / INFO lines to stdout when NCCL_DEBUG is set; stdout carries the one JSON line
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

SEGMENT_BYTES = {"tet4": 308.0, "hex8": 412.0, "quad4": 280.0}  # SURVEY.md 8(d) / BASELINE.md section 2
METRIC = "neutron histories/sec (active batches)"


def build_problem(name, particles_total, n_inactive):
    from macaw_b200 import cstructs, meshgen, nucdata
    if name == "pin_cell_tet":
        mesh = meshgen.pin_cell_tet(26, 25)  # 26*26*25*6 = 101 400 TET4
        lib = nucdata.make_library(9, G=10000, seed=1)
        h = meshgen.PIN_PITCH / 2
        box = (-h, -h, -h, h, h, h)
        et = "tet4"
    elif name == "assembly_quad":
        mesh = meshgen.assembly_quad(17, 9)
        lib = nucdata.make_library(35, G=10000, seed=1)
        L = 17 * meshgen.PIN_PITCH
        box = (0, 0, -1, L, L, 1)
        et = "quad4"
    elif name == "beavrs_assembly":
        # the reference's own mesh (31enr_16_in.e via tests/golden/make_beavrs_mesh.py), 35 synthetic nuclides
        mesh = meshgen.beavrs_assembly_file(os.path.join(ROOT, "tests", "golden", "beavrs_assembly_mesh.npz"))
        lib = nucdata.make_library(35, G=10000, seed=1)
        box = (-0.6, -20.7, -1, 20.7, 0.6, 1)  # examples/beavrs_assembly/settings.xml source box
        et = "quad4"
    elif name == "assembly_hex":
        mesh = meshgen.assembly_hex(17, 9, 4)  # 153*153*4 = 93 636 HEX8 (reference 3-D assembly: 95 504)
        lib = nucdata.make_library(35, G=10000, seed=1)
        L = 17 * meshgen.PIN_PITCH
        box = (0, 0, 0, L, L, 4.0)
        et = "hex8"
    elif name.startswith("core_hex_"):
        # C4: BEAVRS-core-sized synthetic (examples/beavrs_core/run_3d_core_cycle1.i: 3-D core, vacuum boundary):
        # n x n assemblies of 17 x 17 pins, `cells` QUAD cells per pin pitch, extruded to `layers` HEX8 layers
        _, _, n_asm, cells, layers = name.split("_")
        mesh = meshgen.core_hex(int(n_asm), int(cells), int(layers))
        if os.environ.get("MCW_BENCH_MORTON"):
            mesh = meshgen.renumber_elements(mesh)
        lib = nucdata.make_library(35, G=10000, seed=1)
        L = int(n_asm) * 17 * meshgen.PIN_PITCH
        box = (0, 0, 0, L, L, mesh.xyz[:, 2].max())
        et = "hex8"
    elif name.startswith("brick_tet_"):
        n = int(name.split("_")[2])  # C5 scaling sweep: 6 n^3 TET4; brick_tet_<n>_morton = elements in Morton order
        mesh = meshgen.brick_tet(n, bc=meshgen.BC_REFLECT)
        if name.endswith("_morton"):
            mesh = meshgen.renumber_elements(mesh)
        lib = nucdata.make_library(9, G=10000, seed=1)
        L = n * 0.05
        box = (0, 0, 0, L, L, L)
        et = "tet4"
    elif name == "infinite_medium":
        mesh = meshgen.infinite_medium(5)
        lib = nucdata.infinite_medium_library(G=10000)
        box = (-5, -5, -5, 5, 5, 5)
        et = "hex8"
    else:
        raise SystemExit(f"unknown workload {name}")
    s = cstructs.make_settings(particles_total, n_inactive=n_inactive, is_2d=mesh.dim == 2, z_coord=100.0,
                               source_box=box, record_traces=0)
    n_mat = len(lib.materials)
    blocks = list(range(n_mat)) if name != "infinite_medium" else [0, 1]
    mats = [b + 1 for b in blocks]
    if name == "beavrs_assembly":
        role = {"fuel": 1, "gap": 2, "clad": 3, "water": 4}  # OpenMC ids of the synthetic library's four materials
        blocks = sorted(meshgen.BEAVRS_BLOCK_ROLE)
        mats = [role[meshgen.BEAVRS_BLOCK_ROLE[b]] for b in blocks]
    return mesh, lib, blocks, mats, s, et


def declared_config(args, mesh, lib, et, world, particles_per_gpu, n_total):
    """The workload definition, identical in both arms (measured quantities go to `stats`)."""
    return {"workload": args.workload, "mesh_elements": int(mesh.n_elem), "elem_type": et,
            "nuclides": len(lib.nuclides), "grid_points_per_nuclide": int(lib.nuclides[0].energy.size),
            "particles_per_gpu_per_batch": int(particles_per_gpu), "global_particles_per_batch": int(n_total),
            "parallelism": f"particle-sharded x{world}", "tally": f"per-element {args.tally.lower()} flux",
            "seed_through_double": 1, "scaling": args.scaling,
            "parity": "GPU == CPU oracle bit for bit on the deterministic-math build (streams, crossing sequences, banks); "
                      "oracle == reference OpenMC code: seeds / counts / bank ids exactly, energies to 1e-12 (DESIGN.md 2)",
            "l2_policy": "a 256 MiB buffer is written before the timed region; every generation streams > 126 MB "
                         "(L2) of particle state (~0.9 GB per 10^6 histories), so inputs are larger than L2"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_counters(workload):
    """Counters of the tracking kernel that only ncu can give (measured DRAM bytes, FP64 pipe share),
    recorded under profiles/ from a `ncu --set full` capture of this workload; None when not captured."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "track_kernel_counters.json")))
        return d.get(workload)
    except Exception:
        return None


def tally_atomics(estimator):
    """ncu-measured FP64 RED traffic of the tracking kernel with the track-length tally on (profiles/); None for the
    collision estimator, whose scores are one RED per collision inside k_collide."""
    if estimator != "TRACKLENGTH":
        return None
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "tracklength_atomics.json")))
    except Exception:
        return None


def cpu_oracle_rate(problem, source_sites, target_seconds, threads, batch, estimator=0):
    """Histories/s of the CPU oracle on a bounded sample of the same workload."""
    from oracle import bridge
    from macaw_b200 import cstructs
    mesh, lib, blocks, mats, s, et = problem

    def run(n):
        o = bridge.Oracle()
        o.set_mesh(mesh, np.asarray(blocks, np.int32))
        o.set_library(lib, [m - 1 for m in mats])
        s2 = cstructs.Settings.from_buffer_copy(bytes(s))
        s2.n_particles = n
        o.set_settings(s2)
        o.add_tally(1, estimator, 0, [0], None)
        sub = source_sites[:n].copy()
        o.set_source_bank(sub)
        t0 = time.perf_counter()
        o.transport_generation(batch, threads)
        return time.perf_counter() - t0

    n0 = min(len(source_sites), 200 * threads)
    t = run(n0)
    n = int(min(len(source_sites), max(n0, n0 * target_seconds / max(t, 1e-3))))
    t = run(n)
    return n / t, n, t


def xs_lookup_leg(device, points=4 * 1024 * 1024, repeats=5):
    """BASELINE.json's second metric: stand-alone macroscopic cross-section lookup on a BEAVRS-shaped
    synthetic library (91 nuclides, ~3e6 grid points, 170 MB > L2), `points` random (E, material) pairs
    resident in HBM.  Reported as (point, nuclide) lookups per second and as ALGORITHMIC GB/s (SURVEY 8d:
    100 B per (point, nuclide) + 44 B per point).  Only the direct kernel reads every bracket from HBM, so
    only its figure is a roofline fraction; the tile kernel reads each grid window once per 256 points and
    is issue-bound -- its DRAM traffic is the ncu figure in profiles/, quoted next to it."""
    import torch
    from macaw_b200 import cstructs, meshgen, nucdata
    from macaw_b200.context import Context
    lib = nucdata.make_library(91, G=32768, seed=1)
    ctx = Context(device)
    ctx.set_mesh(meshgen.infinite_medium(5))
    ctx.set_library(lib, [0, 1], [0, 1])
    ctx.set_settings(cstructs.make_settings(1000, source_box=(-5, -5, -5, 5, 5, 5)))
    ctx.prepare()
    g = torch.Generator(device="cuda").manual_seed(7)
    lo, hi = np.log(1.1e-5), np.log(1.9e7)
    E = torch.exp(torch.rand(points, generator=g, device="cuda", dtype=torch.float64) * (hi - lo) + lo)
    mat = torch.randint(0, len(lib.materials), (points,), generator=g, device="cuda", dtype=torch.int32)
    out = torch.empty(points * 4, device="cuda", dtype=torch.float64)
    n_nuc = torch.tensor([len(m.nuclides) for m in lib.materials], device="cuda")[mat.long()].sum().item()
    alg_bytes = 100.0 * n_nuc + 44.0 * points
    peak, peak_src = peaks()
    res = {"library": f"{len(lib.nuclides)} nuclides x 32768 points, {lib.bytes() / 1e6:.0f} MB", "points": points,
           "nuclide_lookups": int(n_nuc), "algorithmic_bytes": alg_bytes, "peak": peak, "unit": "GB/s",
           "note": "algorithmic GB/s = (100 B per (point, nuclide) + 44 B per point) / time.  frac is reported for the "
                   "direct kernel only (it reads every bracket itself: an HBM figure); the tile kernels serve 256 "
                   "neighbouring energies from one staged window, so their DRAM traffic is far below the algorithmic "
                   "bytes and their limit is instruction issue -- see measured_dram for the ncu figures"}
    try:
        res["measured_dram"] = json.load(open(os.path.join(ROOT, "profiles", "xs_lookup_counters.json")))
    except Exception:
        res["measured_dram"] = None
    ref = None
    for name, mode in (("direct", 0), ("binned_plus_tile", 2)):
        ctx.xs_lookup_device(points, E.data_ptr(), mat.data_ptr(), out.data_ptr(), 1, mode)  # warm-up
        ms = ctx.xs_lookup_device(points, E.data_ptr(), mat.data_ptr(), out.data_ptr(), repeats, mode)
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        res[name] = {"ms": ms, "algorithmic_gbs": gbs, "nuclide_lookups_per_s": n_nuc / (ms * 1e-3)}
        if mode == 0:
            res[name]["frac"] = gbs / peak
        if ref is None:
            ref = out.clone()
        else:
            res["bit_identical"] = bool(torch.equal(ref.view(torch.int64), out.view(torch.int64)))
    # the tile kernel alone, on arrays the caller already ordered by (material, energy)
    key = mat.double() * 1e9 + E
    order = torch.argsort(key)
    Es, ms_ = E[order].contiguous(), mat[order].contiguous()
    ctx.xs_lookup_device(points, Es.data_ptr(), ms_.data_ptr(), out.data_ptr(), 1, 1)
    ms = ctx.xs_lookup_device(points, Es.data_ptr(), ms_.data_ptr(), out.data_ptr(), repeats, 1)
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    res["tile_presorted"] = {"ms": ms, "algorithmic_gbs": gbs, "nuclide_lookups_per_s": n_nuc / (ms * 1e-3)}
    ctx.close()
    return res


def run_reference(args):
    """--impl reference: the reference's CPU algorithms (oracle restatement; the real MaCaw binary cannot be
    built offline) on all host cores, same workload definition and -- at N = 1 -- the same particle
    count per step; at N > 1 the step is a bounded sample (one GPU's share) of the global batch."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import bridge
    threads = os.cpu_count() or 1
    bridge.build(ref=False)
    per_gpu = args.particles if args.scaling == "weak" else max(1, args.particles // max(world, 1))
    n_total = per_gpu * world if args.scaling == "weak" else args.particles
    mesh, lib, blocks, mats, s, et = build_problem(args.workload, 1000, args.warmup)

    def make(n):
        o = bridge.Oracle()
        o.set_mesh(mesh, np.asarray(blocks, np.int32))
        o.set_library(lib, [m - 1 for m in mats])
        s.n_particles = n
        o.set_settings(s)
        o.add_tally(1, int(args.tally == "TRACKLENGTH"), 0, [0], None)
        o.sample_initial_source()
        return o

    # calibrate: the whole K + W run must end within a few minutes
    o = make(100 * threads)
    t0 = time.perf_counter()
    o.run_generation(1, threads)
    rate = 100 * threads / max(time.perf_counter() - t0, 1e-3)
    budget = 150.0 / (args.steps + args.warmup)
    n = int(max(100 * threads, min(min(n_total, args.particles if args.scaling == "weak" else n_total),
                                   0.8 * rate * budget)))
    o = make(n)
    for b in range(1, args.warmup + 1):
        o.run_generation(b, threads)
    t0 = time.perf_counter()
    for b in range(args.warmup + 1, args.warmup + args.steps + 1):
        res = o.run_generation(b, threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = (f"{n} histories/step x {args.steps} steps of {args.workload} on {threads} host threads "
              f"({'the full per-GPU batch' if n == per_gpu else 'bounded sample of the batch'})")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "histories/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
           "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": declared_config(args, mesh, lib, et, world, per_gpu, n_total),
           "stats": {"particles_per_step_run": n, "keff": res.keff, "host_threads": threads},
           "cpu_baseline": {"value": value, "unit": "histories/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


class Runner:
    """One study on this rank's GPU plus the timing plumbing shared by the bench legs."""

    def __init__(self, args, workload, per_gpu, n_inactive, rank, world, local, group):
        import torch
        from macaw_b200 import study as st
        self.torch, self.rank, self.world, self.local = torch, rank, world, local
        n_total = per_gpu * world
        self.n_total = n_total
        self.problem = build_problem(workload, n_total, n_inactive)
        mesh, lib, blocks, mats, s, et = self.problem
        s.sort_xs_queue = args.sort_xs
        kernel = st.CollisionKernel(blocks=blocks, materials=mats, temperature=293.6,
                                    z_coord=100.0 if mesh.dim == 2 else None)
        tally = st.OpenMCTally(id=1, scores=["flux"], filters=["cell"], estimator=args.tally)
        self.study = st.OpenMCStudy(mesh, lib, kernel, s, [tally], device=local, group=group, rank=rank, nranks=world)
        self.ctx = self.study.ctx
        self.ctx.set_option("sync_every", args.sync_every)
        self.ctx.set_option("profile", 1)
        for kv in args.opt:
            k, v = kv.split("=")
            self.ctx.set_option(k, int(v))
        if args.collide_min >= 0:
            self.ctx.set_option("collide_min", args.collide_min)
        if args.tail_threshold >= 0:
            self.ctx.set_option("tail_threshold", args.tail_threshold)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream(), device=torch.device("cuda", local))

    def barrier(self):
        torch = self.torch
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as td
            td.barrier()
        torch.cuda.synchronize()

    def timed(self, steps, sampler=None):
        """`steps` generations bracketed by barrier + synchronize; device ms on the library's stream."""
        torch = self.torch
        acc = {"seg": 0, "col": 0, "launch": 0, "ms_track": 0.0, "ms_xs": 0.0, "ms_col": 0.0, "ms_bank": 0.0,
               "rounds": 0, "ve": 0, "failed": 0, "dropped": 0}
        with torch.cuda.stream(self.stream):
            self.barrier()
            if sampler:
                sampler.start()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            ev0.record(self.stream)
            for _ in range(steps):
                res = self.study.execute()
                acc["seg"] += res.n_segments
                acc["col"] += res.n_collisions
                acc["launch"] += res.n_kernel_launches
                acc["ms_track"] += res.ms_track
                acc["ms_xs"] += res.ms_xs
                acc["ms_col"] += res.ms_collide
                acc["ms_bank"] += res.ms_bank
                acc["rounds"] += res.n_event_rounds
                acc["ve"] += res.n_vertex_edge_hits
                acc["failed"] += res.n_failed_traces
                acc["dropped"] += res.n_bank_dropped
            ev1.record(self.stream)
            self.barrier()
            wall = time.perf_counter() - t0
        return ev0.elapsed_time(ev1), wall * 1e3, acc, res

    def e2e(self, steps):
        """The same metric through the C ABI with HOST buffers: pinned source bank up, next bank + k back."""
        torch = self.torch
        from macaw_b200.cstructs import SITE_DTYPE
        study, ctx, rank = self.study, self.ctx, self.rank
        n_loc = int(study.work_index[rank + 1] - study.work_index[rank])
        pin_in = torch.empty(n_loc * 96, dtype=torch.uint8).pin_memory()
        pin_out = torch.empty(n_loc * 96, dtype=torch.uint8).pin_memory()
        host_in, host_out = pin_in.numpy().view(SITE_DTYPE), pin_out.numpy().view(SITE_DTYPE)
        host_in[:] = ctx.source_bank()
        with torch.cuda.stream(self.stream):
            self.barrier()
            t1 = time.perf_counter()
            for _ in range(steps):
                ctx.set_source_bank(host_in)          # H2D n*96 B
                r2 = study.execute()
                ctx.source_bank(out=host_out)         # D2H n*96 B (next generation's source) + result struct
                k_host = r2.k_generation
                host_in, host_out = host_out, host_in
            self.barrier()
            return (time.perf_counter() - t1) * 1e3, k_host

    def reduce(self, times, sums):
        torch = self.torch
        t = torch.tensor(times, dtype=torch.float64, device="cuda")
        s = torch.tensor(sums, dtype=torch.float64, device="cuda")
        if self.world > 1:
            import torch.distributed as td
            td.all_reduce(t, op=td.ReduceOp.MAX)
            td.all_reduce(s, op=td.ReduceOp.SUM)
        return [float(x) for x in t.tolist()], [float(x) for x in s.tolist()]

    def close(self):
        self.ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="beavrs_assembly")
    ap.add_argument("--particles", type=int, default=1_000_000,
                    help="histories per batch: per GPU with --scaling weak, in total with --scaling strong")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--tally", default="COLLISION", choices=["COLLISION", "TRACKLENGTH"],
                    help="estimator of the per-element flux tally (TRACKLENGTH scores every segment: the atomic-throughput case)")
    ap.add_argument("--sort-xs", type=int, default=1,
                    help="energy-binned tile refresh of the cross sections (settings.sort_xs_queue): the library bins "
                         "the rounds that are large enough for their nuclide count")
    ap.add_argument("--sync-every", type=int, default=4)
    ap.add_argument("--tail-threshold", type=int, default=-1, help="histories in flight below which tracking runs fused")
    ap.add_argument("--collide-min", type=int, default=-1, help="fused tracking: parked collisions per warp before they are processed")
    ap.add_argument("--opt", action="append", default=[], help="mcw_set_option name=value (tuning knobs; results unchanged)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-xs-leg", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the reference_size and pin-cell legs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from macaw_b200 import lib as mcw_lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    group = None
    if world > 1:
        import torch.distributed as td
        td.init_process_group("nccl", device_id=torch.device("cuda", local))

    per_gpu = args.particles if args.scaling == "weak" else max(1, args.particles // world)
    R = Runner(args, args.workload, per_gpu, args.warmup, rank, world, local, group)
    mesh, lib, blocks, mats, s, et = R.problem
    n_total = R.n_total
    with torch.cuda.stream(R.stream):
        for _ in range(args.warmup):
            R.study.execute()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    flush.fill_(1)  # > 126 MB L2, written
    sampler = ClockSampler(local) if rank == 0 else None  # one nvidia-smi poller per job
    ms_dev, wall_ms, acc, res = R.timed(args.steps, sampler)
    clocks = sampler.stop() if sampler else None
    del flush
    e2e_steps = max(1, min(args.steps, 5))
    e2e_ms, k_host = R.e2e(e2e_steps)
    (ms_dev, wall_ms, e2e_ms), (seg_all, col_all, launches_all, ve_all, failed_all, dropped_all) = R.reduce(
        [ms_dev, wall_ms, e2e_ms], [acc["seg"], acc["col"], acc["launch"], acc["ve"], acc["failed"], acc["dropped"]])
    src_bank = R.ctx.source_bank() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    next_batch = args.warmup + args.steps + e2e_steps + 1
    R.close()

    # ---- the same workload at the reference's own size: 200 000 particles x 5 batches, 0 inactive (all timed)
    ref_size = None
    if not args.no_extra and args.workload == "beavrs_assembly":
        n_ref = 200_000
        for attempt in ("warm-up run", "timed run"):  # two fresh studies: the first pages everything in
            R2 = Runner(args, "beavrs_assembly", max(1, n_ref // world), 0, rank, world, local, group)
            ms2, wall2, acc2, res2 = R2.timed(5)
            n2 = R2.n_total
            R2.close()
        (ms2, wall2), (seg2, launches2) = R.reduce([ms2, wall2], [acc2["seg"], acc2["launch"]])
        ref_size = {"workload": "beavrs_assembly", "source": "examples/beavrs_assembly/settings.xml:4-6",
                    "global_particles_per_batch": n2, "batches": 5, "inactive": 0, "value": n2 * 5 / (ms2 * 1e-3),
                    "unit": "histories/s", "ms_per_batch": ms2 / 5, "wall_ms_per_batch": wall2 / 5,
                    "event_rounds_per_batch": acc2["rounds"] / 5, "gpu_launches": int(launches2), "keff": res2.keff,
                    "scaling": "strong" if world > 1 else "n/a"}

    # ---- BASELINE.json configs[1], the UO2 pin cell (round-1 headline), N = 1 only
    extra = None
    if not args.no_extra and world == 1 and args.workload != "pin_cell_tet":
        R3 = Runner(args, "pin_cell_tet", 1_000_000, 3, rank, world, local, group)
        with torch.cuda.stream(R3.stream):
            for _ in range(3):
                R3.study.execute()
        ms3, wall3, acc3, res3 = R3.timed(5)
        extra = {"pin_cell_tet": {"mesh_elements": int(R3.problem[0].n_elem), "elem_type": "tet4", "nuclides": 9,
                                  "particles_per_batch": 1_000_000, "steps": 5, "warmup": 3,
                                  "value": 1_000_000 * 5 / (ms3 * 1e-3), "unit": "histories/s", "ms_per_step": ms3 / 5,
                                  "segments_per_history": acc3["seg"] / 5e6,
                                  "k_track_algorithmic_gbs": SEGMENT_BYTES["tet4"] * acc3["seg"] / (acc3["ms_track"] * 1e-3) / 1e9,
                                  "k_track_share_of_step": acc3["ms_track"] / ms3, "keff": res3.keff}}
        R3.close()

    if rank == 0:
        value = n_total * args.steps / (ms_dev * 1e-3)
        e2e_value = n_total * e2e_steps / (e2e_ms * 1e-3)
        peak, peak_src = peaks()
        bytes_seg = SEGMENT_BYTES[et]
        achieved = bytes_seg * acc["seg"] / (acc["ms_track"] * 1e-3) / 1e9 if acc["ms_track"] > 0 else None
        ncu = ncu_counters(args.workload) or {}
        cpu = None
        if src_bank is not None:  # reported on rank 0 at N = 1 only
            try:
                threads = os.cpu_count() or 1
                rate, n_s, t_s = cpu_oracle_rate(R.problem, src_bank, 15.0, threads, next_batch, int(args.tally == "TRACKLENGTH"))
                cpu = {"value": rate, "unit": "histories/s", "cores": threads, "kind": "port",
                       "sample": f"{n_s} histories of the same source bank, one generation, {t_s:.1f} s, "
                                 f"{threads} host threads (os.cpu_count)"}
            except Exception as e:  # the oracle is only a reported baseline
                cpu = {"value": None, "unit": "histories/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        out = {
            "metric": METRIC, "value": value, "unit": "histories/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": declared_config(args, mesh, lib, et, world, per_gpu, n_total),
            "stats": {"keff": res.keff, "keff_std": res.keff_std,
                      "segments_per_history": seg_all / (n_total * args.steps),
                      "collisions_per_history": col_all / (n_total * args.steps),
                      "event_rounds_per_batch": acc["rounds"] / args.steps,
                      "vertex_edge_hits": ve_all, "failed_traces": failed_all, "bank_sites_dropped": dropped_all,
                      "wall_ms_per_step": wall_ms / args.steps, "sort_xs_queue": args.sort_xs,
                      "library_loaded": mcw_lib.loaded_path()},
            "e2e": {"value": e2e_value, "unit": "histories/s", "h2d_bytes_per_step": n_total * 96,
                    "d2h_bytes_per_step": n_total * 96 + 8 * world, "steps": e2e_steps, "k_read_back": k_host},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_track", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": ncu.get("dram_bytes_per_launch"),
                         "peak_source": peak_src, "algorithmic_bytes_per_segment": bytes_seg,
                         "segments_per_launch": acc["seg"] / max(acc["rounds"], 1),
                         "avg_launch_ms": acc["ms_track"] / max(acc["rounds"], 1),
                         "kernel_share_of_step": acc["ms_track"] / ms_dev,
                         # what actually limits the kernel (ncu --set full, profiles/): the mesh of this workload
                         # lives in L2 and lane state in registers, so DRAM traffic is far below the byte model
                         "dram_measured_gbs": ncu.get("dram_gbs"), "fp64_pipe_frac": ncu.get("fp64_pipe_frac"),
                         "issue_active_frac": ncu.get("issue_active_frac"), "ncu_source": ncu.get("source"),
                         "note": "achieved/frac follow the SURVEY 8d byte model (state 192 B + element record per segment); "
                                 "measured DRAM traffic is far lower (L2-resident mesh, register-resident lanes): the kernel is "
                                 "FP64-issue / latency bound, see fp64_pipe_frac and dram_measured_gbs",
                         "other_kernels_ms_per_step": {"xs_lookup": acc["ms_xs"] / args.steps,
                                                       "collide+fallback": acc["ms_col"] / args.steps,
                                                       "bank+reduce": acc["ms_bank"] / args.steps}},
            "cpu_baseline": cpu,
            "tally_atomics": tally_atomics(args.tally),
            "reference_size": ref_size,
            "extra": extra,
        }
        if not args.no_xs_leg and world == 1:
            try:
                out["xs_lookup"] = xs_lookup_leg(local)
            except Exception as e:
                out["xs_lookup"] = {"failed": str(e)}
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as td
        td.destroy_process_group()


if __name__ == "__main__":
    main()
