"""One short pass of the hot path for ncu (tools only): structure build, assembly of A and b, `its` CG iterations.
    python tools/prof_step.py --n 4096 --its 12 [--scatter rows|colored]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfel_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--its", type=int, default=12)
ap.add_argument("--scatter", default="rows")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
ctx = capi.Context(0)
mesh = capi.Mesh.square(ctx, 1.0, 1.0, a.n, a.n)
space = capi.Space(mesh, "p1")
space.add_dirichlet_boundary(0.0)
A = capi.Matrix(space, space, scatter=a.scatter)
b = capi.Vector(space)
s = capi.Solver(ctx, method="cg", precond="jacobi", maxits=a.its, rtol=1e-8, check_every=a.its)
for _ in range(a.reps):
    A.clear(); A.assemble("qf5pT", [dict(test=(0, 1), trial=(0, 1)), dict(test=(0, 2), trial=(0, 2))])
    b.clear(); b.assemble("qf5pT", [dict(test=(0, 0))])
    x, rep = A.solve(b, s)
print("prof_step: n=%d nnz=%d its=%d rnorm=%.3e launches=%d" % (a.n, A.nnz, rep["its"], rep["rnorm"], ctx.launches))
