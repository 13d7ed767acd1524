"""SpMV tile-shape sweep on the configs[1] matrix (tools only): GB/s of algorithmic bytes per configuration."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfel_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfgs = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4]
ctx = capi.Context(0)
mesh = capi.Mesh.square(ctx, 1.0, 1.0, n, n)
space = capi.Space(mesh, "p1")
space.add_dirichlet_boundary(0.0)
A = capi.Matrix(space, space)
A.assemble("qf5pT", [dict(test=(0, 1), trial=(0, 1)), dict(test=(0, 2), trial=(0, 2))])
x = np.random.RandomState(0).standard_normal(A.n_rows)
s = capi.Solver(ctx, spmv="vector"); s.set_operator(A); y_ref = s.spmv(x)
ms, nb = s.spmv_bench(30); print("vector   : %.3f ms  %.0f GB/s" % (ms, nb / ms / 1e6)); s.close()
for c in cfgs:
    os.environ["TFELCU_SPMV_CFG"] = str(c)
    s = capi.Solver(ctx, spmv="stream"); s.set_operator(A)
    y = s.spmv(x)
    err = np.abs(y - y_ref).max() / np.abs(y_ref).max()
    ms, nb = s.spmv_bench(50)
    print("stream %d : %.3f ms  %.0f GB/s  (max rel diff vs vector kernel %.1e)" % (c, ms, nb / ms / 1e6, err))
    s.close()
