"""configs[2] probe (tools only): 3-D Laplacian, tetrahedron P2 (new element) on gen_cube_mesh(n), CG + block-Jacobi."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cuda_runner  # noqa: E402
import problems  # noqa: E402
from tfel_b200 import capi  # noqa: E402

ctx = capi.Context(0)
for n in [int(v) for v in sys.argv[1:]] or [32]:
    p = problems.tet("p2", n)
    t0 = time.perf_counter()
    B = cuda_runner.build(ctx, p)
    ctx.synchronize(); t_struct = time.perf_counter() - t0
    for rep in range(2):
        B.matrix.clear(); B.vector.clear()
        t0 = time.perf_counter()
        cuda_runner.assemble(B)
        ctx.synchronize(); t_asm = time.perf_counter() - t0
    for pc, kw in (("jacobi", {}), ("block_jacobi", dict(block_size=4))):
        s = capi.Solver(ctx, method="cg", precond=pc, maxits=20000, rtol=1e-8, check_every=50, **kw)
        x, rep = B.matrix.solve(B.vector, s)
        xyz = B.spaces[0].dof_coordinates()
        err = np.abs(x - (xyz ** 2).sum(1)).max()
        print("tet P2 n=%d: cells %d dofs %d nnz %d | structure %.2f s, assembly %.4f s (%.3g cells/s) | cg+%s: its %d converged %d "
              "solve %.3f s (%.3f ms/it) max|u-|x|^2| %.2e" % (n, B.mesh.n_cells, B.matrix.n_rows, B.matrix.nnz, t_struct, t_asm,
                                                          B.mesh.n_cells / t_asm, pc, rep["its"], rep["converged"],
                                                          rep["t_solve_s"], 1e3 * rep["t_solve_s"] / max(rep["its"], 1), err), flush=True)
        s.close()
    cuda_runner.close(B)
