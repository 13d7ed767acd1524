"""TEST INFRASTRUCTURE: run a tests/problems.py problem through the plain-C oracle.

Produces arrays under the same names oracle/ref_driver.cpp dumps, so that the result can be
compared key by key with tests/golden/*.npz (the reference's own output) and, in the GPU
tests, with what the CUDA library returns.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import oracle as orc  # noqa: E402
from problems import pin_cell  # noqa: E402


def make_mesh(problem):
    kind, n = problem.mesh
    if kind == "square":
        return orc.gen_square_mesh(1.0, 1.0, n, n)
    return orc.gen_cube_mesh(1.0, 1.0, 1.0, n, n, n)


def run(problem, assemble=True):
    dim = problem.dim
    vertices, cells = make_mesh(problem)
    K = cells.shape[0]
    out = {"vertices": vertices, "cells": cells}
    out["cell_volume"], out["jmt"] = orc.geometry(vertices, cells)
    bel, bid, bsd = orc.boundary_submesh(dim, cells)
    out["boundary_cells"], out["boundary_parent_cell"], out["boundary_parent_subdomain"] = bel, bid, bsd
    out["cell_neighbours"] = orc.cell_neighbours(dim, cells)

    ncomp = len(problem.comps)
    prefix = (lambda c: "") if ncomp == 1 and not problem.composite else (lambda c: "c%d_" % c)
    comps = []
    for c, fe_name in enumerate(problem.comps):
        fe = orc.FE_BY_NAME[fe_name]
        dof_map, N, g2l = orc.fes_build(dim, fe, cells)
        comps.append((fe, dof_map, N, g2l))
        out[prefix(c) + "dof_map"] = dof_map
        out[prefix(c) + "g2l"] = g2l
    space = orc.Space(dim, [(fe, dm, N) for fe, dm, N, _ in comps])

    # Dirichlet sets (fes.hpp:100-149): first insert wins
    dir_ids = [dict() for _ in comps]
    for spec in problem.dirichlet:
        c = spec["comp"]
        fe, dof_map, N, g2l = comps[c]
        if spec["where"] == "boundary":
            bcells = bel
        else:  # pin: point submesh at the first vertex of a cell (mesh.hpp:470-476)
            bcells = np.array([[cells[pin_cell(problem.mesh[1]), 0]]], dtype=np.uint32)
            out["pin_cells"] = bcells
            out["pin_parent_cell"] = np.array([pin_cell(problem.mesh[1])], dtype=np.uint32)
            out["pin_parent_subdomain"] = np.zeros(1, dtype=np.uint32)
        dofs = orc.dirichlet_dofs(dim, fe, cells, bcells)
        if callable(spec["value"]):
            x = orc.dof_coordinates(dim, fe, vertices, cells, g2l, dofs)
            vals = spec["value"](x)
        else:
            vals = np.full(dofs.shape[0], float(spec["value"]))
        for d, v in zip(dofs.tolist(), vals.tolist()):
            dir_ids[c].setdefault(d, v)
    flags = np.zeros(space.total, dtype=np.uint8)
    gvals = np.zeros(space.total)
    for c in range(ncomp):
        ids = np.array(sorted(dir_ids[c]), dtype=np.uint32)
        vals = np.array([dir_ids[c][i] for i in ids.tolist()])
        out[prefix(c) + "dirichlet_dof"] = ids
        out[prefix(c) + "dirichlet_value"] = vals
        flags[space.offsets[c] + ids.astype(np.int64)] = 1
        gvals[space.offsets[c] + ids.astype(np.int64)] = vals
    out["dirichlet_flags"] = flags
    if not assemble:
        return out

    # coefficient tables at the quadrature points
    xq = orc.quadrature_points(vertices, cells, problem.quad)
    field_coef = {}
    for name, (c, fn) in problem.fields.items():
        fe, dof_map, N, g2l = comps[c]
        x = orc.dof_coordinates(dim, fe, vertices, cells, g2l, np.arange(N, dtype=np.uint32))
        field_coef[name] = fn(x)
        out["coef_" + name] = field_coef[name]

    def table_of(term):
        tab = None
        for f in term.get("factors", []):
            if f[0] == "x":
                t = f[1].eval(xq) if hasattr(f[1], "eval") else np.full(xq.shape[:2], float(f[1]))
            else:
                c, _ = problem.fields[f[1]]
                fe, dof_map, N, g2l = comps[c]
                t = orc.field_at_quadrature(fe, vertices, cells, dof_map, field_coef[f[1]], f[2], problem.quad)
            tab = t if tab is None else tab * t
        return tab

    def lower(terms):
        return [dict(test=t["test"], trial=t.get("trial", (-1, 0)), c=t.get("c", 1.0), table=table_of(t))
                for t in terms]

    rowptr, colind = orc.pattern(K, space, space, flags)
    val = np.zeros(colind.shape[0])
    # clear(): identity rows (bilinear_form.hpp:159-165)
    for r in np.nonzero(flags)[0].tolist():
        e = rowptr[r] + np.searchsorted(colind[rowptr[r]:rowptr[r + 1]], r)
        val[e] += 1.0
    orc.assemble_bilinear(vertices, cells, space, space, problem.quad, lower(problem.bilinear),
                          flags, problem.composite, rowptr, colind, val)
    b = orc.assemble_linear(vertices, cells, space, problem.quad, lower(problem.linear), problem.composite)
    rhs = b.copy()
    rhs[flags == 1] = gvals[flags == 1]   # bilinear_form.hpp:134-137
    out.update(csr_rowptr=rowptr, csr_colind=colind, csr_val=val, linear_form=b, rhs=rhs)
    return out
