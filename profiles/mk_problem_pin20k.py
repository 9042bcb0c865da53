import sys; sys.path.insert(0,'.')
from tests import problems
from macaw_b200 import problem_io
mesh, lib, bm, s = problems.pin_cell_tet(n_particles=20000)
problem_io.write_problem('gpurun_out/pin_20k.bin', mesh, lib, list(range(len(bm))), [m+1 for m in bm], s, [(1,0,0,problems.ALL_SCORES,problems.EBINS)])
