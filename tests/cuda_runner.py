"""Run a tests/problems.py problem through the CUDA library (C ABI via tfel_b200.capi) and return arrays
under the names oracle_runner.run / the golden fixtures use."""
import numpy as np

from problems import pin_cell
from tfel_b200 import capi


def make_mesh(ctx, problem, generator="device"):
    kind, n = problem.mesh
    if generator == "device":
        return capi.Mesh.square(ctx, 1.0, 1.0, n, n) if kind == "square" else capi.Mesh.cube(ctx, 1.0, 1.0, 1.0, n, n, n)
    import oracle as orc   # host arrays through tfelcu_mesh_create (tests only)
    v, c = orc.gen_square_mesh(1.0, 1.0, n, n) if kind == "square" else orc.gen_cube_mesh(1.0, 1.0, 1.0, n, n, n)
    return capi.Mesh.create(ctx, v, c)


class Built:
    pass


def build(ctx, problem, scatter=None, generator="device"):
    """mesh, spaces with Dirichlet data, matrix, vector (not yet assembled)."""
    B = Built()
    B.problem = problem
    B.mesh = make_mesh(ctx, problem, generator)
    B.spaces = [capi.Space(B.mesh, fe) for fe in problem.comps]
    B.pin = None
    for spec in problem.dirichlet:
        sp = B.spaces[spec["comp"]]
        if spec["where"] == "boundary":
            bcells = None
        else:
            _, cells = B.mesh.download()
            bcells = np.array([[cells[pin_cell(problem.mesh[1]), 0]]], dtype=np.uint32)
            B.pin = bcells
        sp.add_dirichlet_boundary(spec["value"], bcells)
    B.fields = {}
    for name, (c, fn) in problem.fields.items():
        B.fields[name] = (B.spaces[c], fn(B.spaces[c].dof_coordinates()))
    B.matrix = capi.Matrix(B.spaces, B.spaces, scatter=scatter)
    B.vector = capi.Vector(B.spaces, scatter=scatter)
    return B


def lower(terms):
    out = []
    for t in terms:
        d = dict(test=t["test"], c=t.get("c", 1.0))
        if "trial" in t:
            d["trial"] = t["trial"]
        fl = []
        for f in t.get("factors", []):
            fl.append(("x", f[1]) if f[0] == "x" else ("field", f[1], f[2]))
        d["factors"] = fl
        out.append(d)
    return out


def assemble(B):
    p = B.problem
    B.matrix.assemble(p.quad, lower(p.bilinear), B.fields)
    B.vector.assemble(p.quad, lower(p.linear), B.fields)


def close(B):
    B.vector.close(); B.matrix.close()
    for s in B.spaces:
        s.close()
    B.mesh.close()


def run(ctx, problem, scatter=None, generator="device", do_assemble=True):
    B = build(ctx, problem, scatter, generator)
    out = {}
    out["vertices"], out["cells"] = B.mesh.download()
    out["cell_volume"], out["jmt"] = B.mesh.geometry()
    out["boundary_cells"], out["boundary_parent_cell"], out["boundary_parent_subdomain"] = B.mesh.boundary()
    out["cell_neighbours"] = B.mesh.neighbours()
    ncomp = len(problem.comps)
    prefix = (lambda c: "") if ncomp == 1 and not problem.composite else (lambda c: "c%d_" % c)
    for c, sp in enumerate(B.spaces):
        out[prefix(c) + "dof_map"] = sp.dof_map()
        out[prefix(c) + "g2l"] = sp.g2l()
        out[prefix(c) + "dirichlet_dof"], out[prefix(c) + "dirichlet_value"] = sp.dirichlet()
    if B.pin is not None:
        out["pin_cells"] = B.pin
    for name, (sp, coef) in B.fields.items():
        out["coef_" + name] = coef
    if do_assemble:
        assemble(B)
        out["csr_rowptr"], out["csr_colind"], out["csr_val"] = B.matrix.csr()
        out["linear_form"] = B.vector.download()
        out["rhs"] = B.matrix.make_rhs(B.vector)
    else:
        out["csr_rowptr"], out["csr_colind"] = B.matrix.csr(values=False)
    out["_built"] = B
    return out
