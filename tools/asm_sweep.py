"""Assembly of configs[1] (Poisson P1, A and b) timed per strategy and per patch geometry (tools only).
    python tools/asm_sweep.py --n 4096 [--configs rows patch:256:256 patch:256:128 patch:512:256 ...]
Each configuration builds its own forms (the patch plan is part of the one-time structure), runs `reps` timed
clear() + `a += ...` and `b += ...` passes with CUDA events on the context stream and prints one JSON line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from tfel_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--fe", default="p1")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--configs", nargs="*", default=["rows", "tile", "auto"])
a = ap.parse_args()

os.environ["TFELCU_NO_POOL"] = "1"   # every configuration builds its own structure
torch.cuda.set_device(0)
ctx = capi.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
mesh = capi.Mesh.square(ctx, 1.0, 1.0, a.n, a.n)
space = capi.Space(mesh, a.fe)
space.add_dirichlet_boundary(0.0)
STIFF = [dict(test=(0, 1), trial=(0, 1)), dict(test=(0, 2), trial=(0, 2))]
SRC = [dict(test=(0, 0))]
ref = None
for cfg in a.configs:
    for k in list(os.environ):
        if k.startswith("TFELCU_P1_") or k in ("TFELCU_NO_VTAB", "TFELCU_P1STREAM", "TFELCU_NO_UNIFORM"): os.environ.pop(k)
    envs = cfg.split("+")[1:]            # auto+TFELCU_P1_MINB=9: environment knobs of one configuration
    for kv in envs:
        k, v = kv.split("=")
        os.environ[k] = v
    parts = cfg.split("+")[0].split(":")
    scatter = parts[0]
    os.environ.pop("TFELCU_PATCH_ROWS", None); os.environ.pop("TFELCU_PATCH_THREADS", None); os.environ.pop("TFELCU_NO_REC", None)
    if scatter == "rows-norec":   # owner-computes rows without the packed (cell | slots | local dof) records
        scatter = "rows"
        os.environ["TFELCU_NO_REC"] = "1"
    os.environ.pop("TFELCU_ROWS_PER_CTA", None)
    if "@" in scatter:   # rows@64: rows (= threads) per CTA of the owner-computes kernel
        scatter, rpc = scatter.split("@")
        os.environ["TFELCU_ROWS_PER_CTA"] = rpc
    if len(parts) > 1:
        os.environ["TFELCU_PATCH_ROWS"] = parts[1]
    if len(parts) > 2:
        os.environ["TFELCU_PATCH_THREADS"] = parts[2]
    ev = lambda: torch.cuda.Event(enable_timing=True)
    e0, e1 = ev(), ev()
    e0.record(stream)
    A = capi.Matrix(space, space, scatter=None if scatter == "auto" else scatter)
    b = capi.Vector(space, scatter=None if scatter in ("auto", "tile") else scatter)
    tA, tb = [], []
    for rep in range(a.reps + 2):
        s0, s1, s2 = ev(), ev(), ev()
        s0.record(stream)
        A.clear(); A.assemble("qf5pT", STIFF)
        s1.record(stream)
        b.clear(); b.assemble("qf5pT", SRC)
        s2.record(stream)
        s2.synchronize()
        if rep == 0:
            e1.record(stream); e1.synchronize()
            first_ms = e0.elapsed_time(e1)
        if rep >= 2:
            tA.append(s0.elapsed_time(s1)); tb.append(s1.elapsed_time(s2))
    _, _, val = A.csr()
    bv = b.download()
    if ref is None:
        ref = (val, bv)
    dA = float(np.abs(val - ref[0]).max() / np.abs(ref[0]).max()); db = float(np.abs(bv - ref[1]).max() / np.abs(ref[1]).max())
    cells = mesh.n_cells
    print(json.dumps({"config": cfg, "n": a.n, "fe": a.fe, "A_ms": round(float(np.mean(tA)), 4), "A_min_ms": round(float(np.min(tA)), 4),
                      "b_ms": round(float(np.mean(tb)), 4), "Ab_ms": round(float(np.mean(tA) + np.mean(tb)), 4),
                      "Gcells_per_s": round(cells / ((np.mean(tA) + np.mean(tb)) * 1e-3) / 1e9, 2),
                      "first_pass_incl_structure_ms": round(first_ms, 1), "rel_diff_A_vs_first": dA, "rel_diff_b_vs_first": db}), flush=True)
    A.close(); b.close()
