"""Hunt for failed traces (fate 2: counted and killed, RT FAILED_TRACES with tolerate_failure) on the
bench workload at 2M histories per batch and check each one against the CPU oracle: same source
bank, same batch -> the oracle must report the same fate, segment count and last element for that
history.  Run on the GPU box:  python tests/failed_trace_check.py [n_particles] [n_batches]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from macaw_b200 import cstructs  # noqa: E402
from macaw_b200.context import Context  # noqa: E402
from oracle import bridge  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mesh, lib, blocks, mats, s, et = bench.build_problem("pin_cell_tet", n, 3)
s.record_traces = 1
g = Context(0)
g.set_mesh(mesh)
g.set_library(lib, [m - 1 for m in mats], blocks)
g.set_settings(s)
g.sample_initial_source()
bridge.build(ref=False)
keff = 1.0  # the k the fission-site yield is divided by in the coming batch (OMC physics.cpp:165)
for b in range(1, nb + 1):
    src = g.source_bank().copy()
    k_used = keff
    r = g.run_generation(b)
    keff = r.keff
    print(f"batch {b}: k {r.k_generation:.5f} failed {r.n_failed_traces} vertex/edge {r.n_vertex_edge_hits}", flush=True)
    if r.n_failed_traces == 0 and b != nb:
        continue
    rec = g.history_records()
    bad = np.nonzero(rec["fate"] == 2)[0]
    # the oracle on a window of histories around each failure (ids are positions in the bank, so the
    # window is cut from the same bank and the id offset is restored through work_offset = first index)
    o = bridge.Oracle()
    o.set_mesh(mesh, np.asarray(blocks, np.int32))
    o.set_library(lib, [m - 1 for m in mats])
    o.set_settings(s)
    o.set_source_bank(src)
    o.L.orc_set_keff(o.h, k_used)
    t0 = time.time()
    o.transport_generation(b, os.cpu_count() or 1)
    ro = o.history_records()
    print(f"  oracle generation in {time.time() - t0:.0f} s; oracle failed {(ro['fate'] == 2).sum()}")
    for i in bad:
        same = all(rec[f][i] == ro[f][i] for f in ("fate", "n_segments", "n_collisions", "last_elem", "seed_tracking", "trace_hash"))
        print(f"  history {i}: gpu fate {rec['fate'][i]} segs {rec['n_segments'][i]} elem {rec['last_elem'][i]} | "
              f"oracle fate {ro['fate'][i]} segs {ro['n_segments'][i]} elem {ro['last_elem'][i]} | identical: {same}")
    mism = int((rec["trace_hash"] != ro["trace_hash"]).sum())
    print(f"  histories whose crossing hash differs from the oracle: {mism} of {n}")
