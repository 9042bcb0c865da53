"""The C++ host layer (host/: OpenMCStudy / CollisionKernel / OpenMCTally / OpenMCTallyAux over the
C ABI + the mcw_run driver).  CPU: it builds, reads the problem file the Python side wrote, rejects
bad parameters with the reference's messages and fails loudly without a GPU.  GPU: its k per batch
and its tally read-back equal the oracle's on the same problem file."""
import json
import os
import subprocess

import numpy as np
import pytest

from macaw_b200 import cstructs, problem_io
from tests import problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host")
TALLIES = [(1, 0, 0, problems.ALL_SCORES, problems.EBINS), (2, 0, 1, [0, 3], None)]


def _build():
    from macaw_b200 import build
    build.build()
    r = subprocess.run(["make", "-C", HOST, "NCCL=1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return os.path.join(HOST, "mcw_run")


def _write(tmp_path, name, **kw):
    mesh, lib, bm, s = problems.PROBLEMS[name](**kw)
    path = str(tmp_path / f"{name}.bin")
    problem_io.write_problem(path, mesh, lib, list(range(len(bm))), [m + 1 for m in bm], s, TALLIES)
    return path, (mesh, lib, bm, s)


def test_problem_file_round_trip(tmp_path):
    path, (mesh, lib, bm, s) = _write(tmp_path, "pin_cell_tet", n_particles=100)
    r = problem_io.read_records(path)
    assert np.array_equal(r["mesh.conn"].reshape(mesh.conn.shape), mesh.conn)
    assert r["mesh.xyz"].tobytes() == np.ascontiguousarray(mesh.xyz, np.float64).tobytes()
    assert r["lib.counts"].tolist() == [len(lib.nuclides), len(lib.materials), lib.n_log_bins]
    assert r["nuc3.xs"].tobytes() == np.ascontiguousarray(lib.nuclides[3].xs, np.float64).tobytes()
    assert r["settings.ints"][0] == 100 and r["kernel.materials"].tolist() == [1, 2, 3, 4]


def test_cpp_host_builds_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build()
    path, _ = _write(tmp_path, "infinite_medium", n_particles=100)
    r = subprocess.run([exe, path, "--batches", "1"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 1 and "no CPU fallback" in r.stderr and r.stdout == ""
    r = subprocess.run([exe, str(tmp_path / "missing.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open problem file" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["infinite_medium", "pin_cell_tet", "assembly_quad", "pin_cell_tet_iso"])
def test_cpp_host_matches_oracle(oracle_bridge, tmp_path, name):
    exe = _build()
    path, (mesh, lib, bm, s) = _write(tmp_path, name, n_particles=3000)
    r = subprocess.run([exe, path, "--batches", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [json.loads(x) for x in r.stdout.splitlines()]
    o = problems.configure(oracle_bridge.Oracle(), mesh, lib, bm, s, True, False)
    o.sample_initial_source()
    for b in (1, 2, 3):
        ro, rc = o.run_generation(b), lines[b - 1]
        assert rc["n_segments"] == ro.n_segments and rc["n_collisions"] == ro.n_collisions
        assert rc["n_bank_sites"] == ro.n_bank_sites
        for f in ("k_generation", "keff", "k_collision", "k_absorption", "k_tracklength"):
            assert abs(rc[f] - getattr(ro, f)) <= 1e-11 * abs(getattr(ro, f)), f
    _, so, _, _ = o.get_tally(1)
    ns = mesh.n_elem
    flux = so.reshape(ns, len(problems.EBINS) - 1, len(problems.ALL_SCORES))[:, :, 0].sum(axis=1)
    t1 = [t for t in lines[-1]["tallies"] if t["id"] == 1][0]
    assert t1["n_spatial"] == ns
    assert abs(t1["sum"] - flux.sum()) <= 1e-10 * abs(flux.sum())
    assert abs(t1["first"] - flux[0]) <= 1e-10 * max(abs(flux[0]), 1e-300)
