"""Tet P2 assembly for ncu (tools only): python tools/prof_tet.py [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cuda_runner  # noqa: E402
import problems  # noqa: E402
from tfel_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = capi.Context(0)
B = cuda_runner.build(ctx, problems.tet("p2", n))
for _ in range(2):
    B.matrix.clear(); B.vector.clear()
    cuda_runner.assemble(B)
ctx.synchronize()
print("prof_tet: n=%d cells=%d dofs=%d nnz=%d" % (n, B.mesh.n_cells, B.matrix.n_rows, B.matrix.nnz))
