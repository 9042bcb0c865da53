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
        self.nuclides = [str(x) for x in nuclides]  # "total" | "all" | names (OpenMCTally.C:56-58)
        self.id, self.scores, self.filters = int(id), list(scores), list(filters)
        self.estimator = cstructs.ESTIMATOR[est.lower()]
        self.energy_bins = list(energy_bins) if energy_bins is not None else None
        self.granularity = (cstructs.GRANULARITY["cell"] if "cell" in filters else
                            cstructs.GRANULARITY["universe"] if "universe" in filters else cstructs.GRANULARITY["global"])


    def nuclide_indices(self, library):
        """What each nuclide bin of results_ scores (OpenMCTally.C:188-204): a list of names -> the list, then total (-1);
        "total" -> the single total bin; "all" -> bin i = nuclide i, bin n = total and one more bin that is never scored
        (-2): the reference appends to the default nuclides_ = {-1} without clearing it, so its results_ has n + 2
        nuclide bins.  self.ref_nuclides is the reference's nuclides_ vector itself (what OpenMCTallyAux.C:93-106 looks
        at).  Unknown names raise like OMC tally.cpp:552-554."""
        if self.nuclides[0] == "all":
            n = len(library.nuclides)
            self.ref_nuclides = [-1] + list(range(n)) + [-1]
            return list(range(n)) + [-1, -2]
        self.ref_nuclides = [-1]
        if self.nuclides[0] == "total":
            return [-1]
        index = {n.name: i for i, n in enumerate(library.nuclides)}
        out = []
        for name in self.nuclides:
            if name == "total":
                out.append(-1)
            elif name in index:
                out.append(index[name])
            else:
                raise ValueError(f"Could not find the nuclide {name} specified in tally {self.id} in any material")
        self.ref_nuclides = out + [-1]
        return out + [-1]


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
            t.nuclide_bins = t.nuclide_indices(library)
            if t.nuclide_bins != [-1]:
                self.ctx.set_tally_nuclides(t.id, t.nuclide_bins)
        self.settings = settings
        self.library = library
        self.work_index = dist.calculate_work(settings.n_particles, nranks)
        self.batch = 0
        self.ctx.prepare()
        self.ctx.sample_initial_source()  # openmc_simulation_init -> initialize_source (OpenMCStudy.C:147)
        self._views = None
        self._ext_stream = None
        if nranks > 1:
            import torch
            dist.warm_up(group, torch.device("cuda", torch.cuda.current_device()))

    def set_temperature(self, T):
        """The coupled temperature changed (a new time step of the multiphysics solve).  With several temperatures in the
        library the new field re-prepares the problem, which starts the run over: this rank's source bank is carried
        across; k statistics and tally sums restart with the new field (include/macaw_b200.h mcw_set_temperature)."""
        bank = self.ctx.source_bank()
        self.ctx.set_temperature(T)
        if len(bank):
            self.ctx.set_source_bank(bank)

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

    def __init__(self, study: OpenMCStudy, tally_id, score, tally: OpenMCTally = None, granularity="cell", nuclide=None):
        if granularity == "global":
            raise ValueError("global granularity is not supported")  # OpenMCTallyAux.C errors
        self.study, self.tally_id, self.score, self.granularity = study, tally_id, score, granularity
        self._tally = tally
        self.nuclide = nuclide if nuclide is not None else "all"  # OpenMCTallyAux.C:61-62

    def score_index(self):
        """OpenMCTallyAux.C:93-119.  The tally's total bin for nuclide "all" (the last bin), else the named nuclide's bin.
        The reference combines the two as score_bin * n_nuclide_bins + nuclide_bin although OpenMC stores
        nuclide_bin * n_scores + score_bin (tally_scoring.cpp:2400); the two agree when either count is 1.  Kept as the
        reference has it (drop-in), flagged here."""
        scores = self._tally.scores if self._tally is not None else None
        score_bin = scores.index(self.score) if scores else 0
        bins = getattr(self._tally, "ref_nuclides", [-1]) if self._tally is not None else [-1]
        if bins[0] != -1:
            if self.nuclide == "all":
                nuc_bin = len(bins) - 1
            else:
                names = [n.name for n in self.study.library.nuclides]
                nuc_bin = bins.index(names.index(self.nuclide))
        else:
            nuc_bin = 0  # also what a `nuclides = all` tally gets: its nuclides_ starts with the default -1
        return score_bin * len(bins) + nuc_bin

    def compute(self):
        s, _, _ = self.study.ctx.get_tally(self.tally_id)
        return s[:, :, self.score_index()].sum(axis=1)

    def compute_value(self, elem):
        return float(self.compute()[elem])
