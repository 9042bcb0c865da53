"""tests/golden/beavrs_assembly_mesh.npz: the reference's own BEAVRS assembly mesh
(examples/beavrs_assembly/31enr_16_in.e, 23 876 QUAD4, blocks 10004 / 10006 / 10010 / 10014 / 11006,
reflecting side sets 1-4 as in run_assembly.i:48-53) converted with macaw_b200/exodus.py, so the GPU
box -- which has no /root/reference -- can run BASELINE.json configs[2] on the real mesh.

  python tests/golden/make_beavrs_mesh.py [/root/reference]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from macaw_b200 import exodus  # noqa: E402

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
m = exodus.read_exodus(os.path.join(ref, "examples", "beavrs_assembly", "31enr_16_in.e"), reflect=(1, 2, 3, 4))
out = os.path.join(ROOT, "tests", "golden", "beavrs_assembly_mesh.npz")
np.savez_compressed(out, xy=m.xyz[:, :2].copy(), conn=m.conn.astype(np.int32), block_id=m.block_id,
                    side_bnd_id=m.side_bnd_id.astype(np.int8), bnd_kind=m.bnd_kind)
print(out, os.path.getsize(out), "bytes;", m.n_elem, "elements")
