"""Host-side mirror of the reference's MOOSE objects for the hot path, over the C ABI.

    OpenMCStudy      <- src/userobjects/OpenMCStudy.C   (one execute() = one batch = one generation)
    CollisionKernel  <- src/raykernels/CollisionKernel.C (configuration carrier: blocks -> materials,
                        temperature, z_coord; its per-segment onSegment() runs inside the GPU kernels)
    OpenMCTally      <- src/userobjects/OpenMCTally.C    (filters / scores / estimator)
    OpenMCTallyAux   <- src/auxkernels/OpenMCTallyAux.C  (per-element read-back of the SUM result)

Parameter names and error behaviour follow the reference's validParams(); see INTEGRATION.md for the
C++ adapter a MOOSE build would compile instead of this file."""
from __future__ import annotations

import numpy as np

from . import cstructs, dist
from .context import Context


class CollisionKernel:
    """params: temperature, blocks, materials (OpenMC material *ids*, 1-based), z_coord, verbose
    (CollisionKernel.C:24-42)."""

    def __init__(self, blocks, materials, temperature=293.6, z_coord=None, verbose=False):
        if len(blocks) != len(materials):
            raise ValueError("The blocks parameter must be the same size as the materials parameter.")  # :66-67
        self.blocks = [int(b) for b in blocks]
        self.materials = [int(m) for m in materials]
        self.temperature, self.z_coord, self.verbose = temperature, z_coord, verbose

    def material_indices(self):
        return [m - 1 for m in self.materials]  # CollisionKernel.C:110


class OpenMCTally:
    """params: id, particle_type, estimator, scores, filters, nuclides, energy_bins, cell_bins,
    block_bins (OpenMCTally.C:30-65).  Filters are sorted universe < cell < energy (OpenMCTally.h:71-88)."""

    def __init__(self, id, scores, filters, estimator="COLLISION", particle_type="neutron", energy_bins=None,
                 nuclides=("total",)):
        if particle_type != "neutron":
            raise ValueError("Only neutrons are currently supported.")  # :87-88
        est = estimator.upper()
        if est not in ("COLLISION", "TRACKLENGTH"):
            raise ValueError("Only collision estimator currently supported")  # :89-90 (+ tracklength extension)
        if "energy" in filters and energy_bins is None:
            raise ValueError("The energy bins parameter should be supplied if an energy filter is requested.")
        for f in filters:
            if f not in ("universe", "cell", "energy", "particle"):
                raise ValueError("Unrecognized filter")  # :173-174
        if tuple(nuclides) != ("total",):
            raise NotImplementedError("per-nuclide tally bins are not on the GPU path yet")
        self.id, self.scores, self.filters = int(id), list(scores), list(filters)
        self.estimator = cstructs.ESTIMATOR[est.lower()]
        self.energy_bins = list(energy_bins) if energy_bins is not None else None
        self.granularity = (cstructs.GRANULARITY["cell"] if "cell" in filters else
                            cstructs.GRANULARITY["universe"] if "universe" in filters else cstructs.GRANULARITY["global"])


class OpenMCStudy:
    """RayTracingStudy-derived transport study (OpenMCStudy.C:78-204).  Construction forwards the
    configuration once; execute() replaces generateRays + the TraceRay loop + postExecuteStudy."""

    def __init__(self, mesh, library, kernel: CollisionKernel, settings, tallies=(), device=0, group=None,
                 rank=0, nranks=1, verbose=False):
        self.rank, self.nranks, self.group, self.verbose = rank, nranks, group, verbose
        self.ctx = Context(device, rank, nranks)
        if mesh.dim == 2 and kernel.z_coord is None:
            raise ValueError("The z_coord parameter must be specified for 2D calculations")  # CollisionKernel.C:60-61
        settings.is_2d = int(mesh.dim == 2)
        if kernel.z_coord is not None:
            settings.z_coord = float(kernel.z_coord)
        self.ctx.set_mesh(mesh)
        self.ctx.set_library(library, kernel.material_indices(), kernel.blocks)
        self.ctx.set_temperature(kernel.temperature)
        self.ctx.set_settings(settings)
        for t in tallies:
            sc = [cstructs.SCORE[s] for s in t.scores]
            self.ctx.add_tally(t.id, t.estimator, t.granularity, sc, t.energy_bins if "energy" in t.filters else None)
        self.settings = settings
        self.work_index = dist.calculate_work(settings.n_particles, nranks)
        self.batch = 0
        self.ctx.prepare()
        self.ctx.sample_initial_source()  # openmc_simulation_init -> initialize_source (OpenMCStudy.C:147)
        self._views = None
        self._ext_stream = None
        if nranks > 1:
            import torch
            dist.warm_up(group, torch.device("cuda", torch.cuda.current_device()))

    # -- one Transient time step (OpenMCStudy.C:366-402)
    def execute(self):
        self.batch += 1
        if self.nranks == 1:
            res = self.ctx.run_generation(self.batch)  # transport + synchronize_bank + finalize on one GPU
        else:
            self.ctx.transport_generation(self.batch)  # generateRays() + ParallelStudy::execute()
            res = self._synchronize_banks_and_reduce(self.batch)  # postExecuteStudy()
        if self.verbose and self.rank == 0:
            print(f"Keff {res.keff} ({res.keff_std}) Generation: {res.k_generation}")  # OpenMCStudy.C:399-401
        return res

    def _synchronize_banks_and_reduce(self, b):
        import torch
        dev = torch.device("cuda", torch.cuda.current_device())
        if self._ext_stream is None:
            self._ext_stream = torch.cuda.ExternalStream(self.ctx.stream(), device=dev)
        # collectives and copies are issued with the library's stream as torch's current stream, so they are
        # ordered with its kernels by the stream itself and no full-stream host synchronisation is needed
        with torch.cuda.stream(self._ext_stream):
            return dist.synchronize_banks(DeviceEngine(self.ctx, dev), b, self.rank, self.nranks, self.work_index,
                                          self.group, dev)


class DeviceEngine:
    """dist.synchronize_banks hooks over the C ABI's raw device views (zero-copy torch tensors, so
    NCCL reads and writes the library's own HBM buffers)."""

    def __init__(self, ctx: Context, device):
        self.ctx, self.device = ctx, device
        for name in ("bank_count", "set_bank_window", "sample_bank", "finish_sampling", "set_source_count"):
            setattr(self, name, getattr(ctx, name))

    def reduction_tensor(self):
        p, n = self.ctx.reduction_device()
        return dist.device_tensor(p, n, "f8", self.device)

    def sampled_tensor(self, n):
        p, _ = self.ctx.sampled_device()
        return dist.device_tensor(p, max(n, 1) * dist.SITE_BYTES, "u8", self.device)

    def source_tensor(self):
        p, cap = self.ctx.source_device()
        return dist.device_tensor(p, cap * dist.SITE_BYTES, "u8", self.device)

    def finalize_generation(self, batch):
        # the exchange and the all-reduce were issued on (or waited for by) the library's own stream
        return self.ctx.finalize_generation(batch)


class OpenMCTallyAux:
    """AuxKernel read-back (OpenMCTallyAux.C:70-215): per element, results_(filter_index, score, SUM)
    summed over the energy bins -- the raw sum over batches, not the mean."""

    def __init__(self, study: OpenMCStudy, tally_id, score, tally: OpenMCTally = None, granularity="cell"):
        if granularity == "global":
            raise ValueError("global granularity is not supported")  # OpenMCTallyAux.C errors
        self.study, self.tally_id, self.score, self.granularity = study, tally_id, score, granularity
        self._tally = tally

    def compute(self):
        s, _, _ = self.study.ctx.get_tally(self.tally_id)
        scores = self._tally.scores if self._tally is not None else None
        j = scores.index(self.score) if scores else 0
        return s[:, :, j].sum(axis=1)

    def compute_value(self, elem):
        return float(self.compute()[elem])
