"""configs[3] probe (tools only): Stokes P2-P2-P1 on gen_square_mesh(n), assembly time and GMRES + ILU(0) behaviour."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cuda_runner  # noqa: E402
import problems  # noqa: E402
from tfel_b200 import capi  # noqa: E402

def _opt(name, default):
    for a in sys.argv[1:]:
        if a.startswith("--%s=" % name):
            return type(default)(a.split("=", 1)[1])
    return default


ILUFILL = _opt("ilufill", 2)       # the reference programs pass ilufill = 2 (test/stokes_2d_p2_p1.cpp:112)
RESTART = _opt("restart", 1000)
MAXITS = _opt("maxits", 6000)
ASSEMBLY_ONLY = "--assembly-only" in sys.argv   # full-size runs (n = 2048: 8.4 M cells, 37.8 M dofs, 1.1e9 stored entries)
ctx = capi.Context(0)
for n in [int(v) for v in sys.argv[1:] if not v.startswith("--")] or [64, 128]:
    p = problems.stokes(n)
    t0 = time.perf_counter()
    B = cuda_runner.build(ctx, p)
    ctx.synchronize(); t_struct = time.perf_counter() - t0
    for rep in range(2):
        B.matrix.clear(); B.vector.clear()
        t0 = time.perf_counter()
        cuda_runner.assemble(B)
        ctx.synchronize(); t_asm = time.perf_counter() - t0
    if ASSEMBLY_ONLY:
        if "--spmv-check" in sys.argv:   # the two SpMV kernels on the full-size operator (indices beyond 2^30 entries)
            import numpy as np
            xr = np.random.RandomState(3).standard_normal(B.matrix.n_rows)
            ys = []
            for kern in ("vector", "stream"):
                sk = capi.Solver(ctx, method="gmres", precond="none", spmv=kern); sk.set_operator(B.matrix)
                ys.append(sk.spmv(xr)); sk.close()
            print("stokes n=%d: spmv vector vs stream max rel diff %.3e" % (n, np.abs(ys[0] - ys[1]).max() / np.abs(ys[0]).max()), flush=True)
        print("stokes n=%d: cells %d dofs %d nnz %d | structure %.3f s, assembly (A and b, 2nd pass) %.4f s (%.3g cells/s)"
              % (n, B.mesh.n_cells, B.matrix.n_rows, B.matrix.nnz, t_struct, t_asm, B.mesh.n_cells / t_asm), flush=True)
        cuda_runner.close(B)
        continue
    s = capi.Solver(ctx, method="gmres", precond="ilu", ilufill=ILUFILL, maxits=MAXITS, restart=RESTART, rtol=1e-8)
    t0 = time.perf_counter()
    x, rep = B.matrix.solve(B.vector, s)
    t_solve = time.perf_counter() - t0
    print("stokes n=%d: cells %d dofs %d nnz %d | structure %.3f s, assembly %.4f s (%.3g cells/s) | gmres(%d)+ilu(%d): its %d "
          "converged %d rnorm/bnorm %.2e | ilu nnz %d, setup %.3f s (precond %.3f s, symbolic %.3f s), solve %.3f s (%.2f ms/it), launches %d"
          % (n, B.mesh.n_cells, B.matrix.n_rows, B.matrix.nnz, t_struct, t_asm, B.mesh.n_cells / t_asm, rep["restart_used"],
             rep["ilu_levels"], rep["its"], rep["converged"], rep["rnorm"] / max(rep["bnorm"], 1e-300), rep["ilu_nnz"],
             rep["t_setup_s"], rep["t_precond_setup_s"], rep["t_symbolic_s"], rep["t_solve_s"],
             1e3 * rep["t_solve_s"] / max(rep["its"], 1), rep["n_launches"]), flush=True)
    s.close(); cuda_runner.close(B)
