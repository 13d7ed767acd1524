import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfel_b200 import capi
os.environ["TFELCU_NO_POOL"] = "1"
ctx = capi.Context(0)
STIFF = [dict(test=(0, 1), trial=(0, 1)), dict(test=(0, 2), trial=(0, 2))]
for n in (1024, 4096):
    mesh = capi.Mesh.square(ctx, 1.0, 1.0, n, n)
    space = capi.Space(mesh, "p1"); space.add_dirichlet_boundary(0.0)
    for knob in (None, "TFELCU_NO_VTAB", None, "TFELCU_NO_UNIFORM", "TFELCU_NO_P1ROWS"):
        for k in ("TFELCU_NO_VTAB", "TFELCU_NO_UNIFORM", "TFELCU_NO_P1ROWS"): os.environ.pop(k, None)
        if knob: os.environ[knob] = "1"
        ctx.synchronize(); t0 = time.perf_counter()
        A = capi.Matrix(space, space); b = capi.Vector(space)
        ctx.synchronize(); t1 = time.perf_counter()
        A.assemble("qf5pT", STIFF); b.assemble("qf5pT", [dict(test=(0, 0))]); b.download()
        ctx.synchronize(); t2 = time.perf_counter()
        A.clear(); A.assemble("qf5pT", STIFF); b.clear(); b.assemble("qf5pT", [dict(test=(0, 0))]); b.download()
        ctx.synchronize(); t3 = time.perf_counter()
        print("n=%d knob=%s: Matrix() %.1f ms, first += (plan) %.1f ms, second += %.1f ms" % (n, knob, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
        A.close(); b.close()
    space.close(); mesh.close()
