"""SpMV back to back at the row counts one rank holds on 1 / 4 / 8 GPUs of configs[1], for every tile configuration of
the streamed kernel (TFELCU_SPMV_CFG): python tools/spmv_sizes.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfel_b200 import capi
os.environ["TFELCU_NO_POOL"] = "1"
ctx = capi.Context(0)
STIFF = [dict(test=(0, 1), trial=(0, 1)), dict(test=(0, 2), trial=(0, 2))]
for n in (1448, 2048, 4096):
    mesh = capi.Mesh.square(ctx, 1.0, 1.0, n, n)
    space = capi.Space(mesh, "p1"); space.add_dirichlet_boundary(0.0)
    A = capi.Matrix(space, space); A.assemble("qf5pT", STIFF)
    for cfg in (0, 1, 2, 3, 4):
        os.environ["TFELCU_SPMV_CFG"] = str(cfg)
        s = capi.Solver(ctx, method="cg", precond="jacobi", maxits=10, rtol=1e-8)
        s.set_operator(A)
        ms, nbytes = s.spmv_bench(200)
        print("n=%d rows=%d cfg=%d: %.1f us per product back to back, %.2f TB/s algorithmic" % (n, A.n_local, cfg, ms * 1e3, nbytes / ms / 1e9), flush=True)
        s.close()
    A.close(); space.close(); mesh.close()
