"""GPU: the CUDA path (through the C ABI) against the committed reference fixtures (tests/golden) and the
plain-C oracle on the same inputs. Integer structure bit-exact, values to 1e-12 (tests/parity.py),
solutions to the solver rtol 1e-8 in relative L2 (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

import cuda_runner
import oracle as orc
import oracle_runner
import problems
from parity import assert_int_equal, assert_values_close, matrix_entries_close

pytestmark = pytest.mark.gpu

SKIP_KEYS = ("meta_json", "solution", "pin_parent_cell", "pin_parent_subdomain", "dirichlet_flags")


@pytest.fixture(scope="module")
def ctx():
    from tfel_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def compare(got, ref, rowptr_key="csr_rowptr"):
    for key, r in ref.items():
        if key in SKIP_KEYS or key.startswith("_"):
            continue
        assert key in got, key
        if r.dtype.kind in "ui":
            assert_int_equal(got[key], r, key)
        elif key == "csr_val":
            matrix_entries_close(ref[rowptr_key], got[key], r, 1e-12)
        else:
            assert_values_close(got[key], r, 1e-12, key)


@pytest.mark.parametrize("scatter", ["rows", "colored", "patch"])
@pytest.mark.parametrize("problem", problems.GOLDEN_CASES, ids=lambda p: p.golden)
def test_cuda_matches_reference_fixture(ctx, problem, scatter, golden_dir):
    gold = dict(np.load(os.path.join(golden_dir, problem.golden + ".npz")))
    got = cuda_runner.run(ctx, problem, scatter=scatter)
    try:
        compare(got, gold)
    finally:
        cuda_runner.close(got["_built"])


@pytest.mark.parametrize("problem", [problems.poisson("p1", 57), problems.poisson("p2", 23), problems.poissong("p2", 19),
                                     problems.tet("p1", 7), problems.tet("p1b", 5), problems.tet("p2", 5),
                                     problems.stokes(13), problems.navier_stokes(9), problems.laplace2(11)],
                         ids=lambda p: "%s_n%d" % (p.name, p.mesh[1]))
def test_cuda_matches_oracle(ctx, problem):
    ref = oracle_runner.run(problem)
    got = cuda_runner.run(ctx, problem)
    try:
        compare(got, ref)
    finally:
        cuda_runner.close(got["_built"])


@pytest.mark.parametrize("problem", [problems.poisson("p1", 57), problems.poisson("p1", 130), problems.poisson("p2", 23),
                                     problems.poissong("p2", 19), problems.tet("p1", 7), problems.tet("p1b", 5),
                                     problems.tet("p2", 5), problems.stokes(13), problems.navier_stokes(9),
                                     problems.laplace2(11)],
                         ids=lambda p: "%s_n%d" % (p.name, p.mesh[1]))
def test_patch_strategy_matches_oracle_and_rows(ctx, problem):
    """TFELCU_SCATTER_PATCH (rows grouped into quadtree / octree leaves, element matrices once per patch in shared
    memory) on every element and form kind, composite forms included: against the oracle to 1e-12, and against the
    owner-computes rows -- same cells per row in the same order, so equal to the last bits."""
    ref = oracle_runner.run(problem)
    got = cuda_runner.run(ctx, problem, scatter="patch")
    rows = cuda_runner.run(ctx, problem, scatter="rows")
    try:
        compare(got, ref)
        scale = np.abs(rows["csr_val"]).max()
        assert np.abs(got["csr_val"] - rows["csr_val"]).max() <= 1e-14 * scale
        assert np.abs(got["linear_form"] - rows["linear_form"]).max() <= 1e-14 * max(np.abs(rows["linear_form"]).max(), 1e-300)
        again = cuda_runner.run(ctx, problem, scatter="patch")
        assert np.array_equal(again["csr_val"], got["csr_val"]) and np.array_equal(again["linear_form"], got["linear_form"])
        cuda_runner.close(again["_built"])
    finally:
        cuda_runner.close(got["_built"]); cuda_runner.close(rows["_built"])


def test_host_mesh_upload_equals_device_generator(ctx):
    p = problems.poisson("p2", 9)
    a = cuda_runner.run(ctx, p, generator="device")
    b = cuda_runner.run(ctx, p, generator="host")
    try:
        for k in ("vertices", "cells", "dof_map", "csr_rowptr", "csr_colind", "csr_val", "linear_form"):
            assert np.array_equal(a[k], b[k]), k
    finally:
        cuda_runner.close(a["_built"]); cuda_runner.close(b["_built"])


def test_assembly_is_bitwise_reproducible(ctx):
    p = problems.stokes(10)
    runs = []
    for scatter in ("rows", "colored"):
        for _ in range(2):
            got = cuda_runner.run(ctx, p, scatter=scatter)
            runs.append(got["csr_val"].copy())
            cuda_runner.close(got["_built"])
    assert np.array_equal(runs[0], runs[1])
    assert np.array_equal(runs[2], runs[3])
    # the two strategies add the cells of a row in different orders (ascending cell id / colour by colour):
    # equal to rounding, not bitwise
    assert np.abs(runs[0] - runs[2]).max() <= 1e-13 * np.abs(runs[0]).max()


@pytest.mark.parametrize("scatter", [None, "rows", "patch"])
def test_accumulate_twice_and_clear(ctx, scatter):
    """a += x; a += x doubles the entries but keeps identity rows; clear() restores them."""
    p = problems.poisson("p1", 12)
    B = cuda_runner.build(ctx, p, scatter=scatter)
    try:
        cuda_runner.assemble(B)
        _, _, v1 = B.matrix.csr()
        B.matrix.assemble(p.quad, cuda_runner.lower(p.bilinear))
        rp, ci, v2 = B.matrix.csr()
        dofs, _ = B.spaces[0].dirichlet()
        diag_pos = np.array([rp[r] for r in dofs])
        is_dir = np.zeros(v1.shape[0], dtype=bool); is_dir[diag_pos] = True
        assert np.array_equal(v2[~is_dir], 2.0 * v1[~is_dir])
        assert np.all(v2[is_dir] == 1.0)
        B.matrix.clear()
        _, _, v0 = B.matrix.csr()
        assert np.all(v0[is_dir] == 1.0) and np.all(v0[~is_dir] == 0.0)
    finally:
        cuda_runner.close(B)


def test_spmv_kernels_match_oracle(ctx):
    from tfel_b200 import capi
    ref = oracle_runner.run(problems.stokes(12))
    a = (ref["csr_rowptr"], ref["csr_colind"], ref["csr_val"])
    x = np.random.RandomState(1).standard_normal(a[0].shape[0] - 1)
    y_ref = orc.spmv(*a, x)
    for kern in ("vector", "stream"):
        s = capi.Solver(ctx, spmv=kern)
        s.set_operator(a)
        y = s.spmv(x)
        s.close()
        assert_values_close(y, y_ref, 1e-14, kern)


def test_spmv_empty_and_ragged_rows(ctx):
    from tfel_b200 import capi
    rs = np.random.RandomState(2)
    n = 5000
    lens = rs.randint(0, 40, n); lens[::7] = 0; lens[123] = 900
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    colind = np.concatenate([np.sort(rs.choice(n, l, replace=False)) for l in lens]).astype(np.int32)
    val = rs.standard_normal(colind.shape[0]); x = rs.standard_normal(n)
    y_ref = orc.spmv(rowptr, colind, val, x)
    for kern in ("vector", "stream"):
        s = capi.Solver(ctx, spmv=kern)
        s.set_operator((rowptr, colind, val))
        y = s.spmv(x)
        s.close()
        assert_values_close(y, y_ref, 1e-13, kern)


@pytest.mark.parametrize("precond", ["none", "jacobi", "block_jacobi", "ilu0"])
def test_cg_solution_parity(ctx, precond):
    """config 1/2 shape: Poisson P1, CG; solution within rtol 1e-8 (relative L2) of the oracle's."""
    from tfel_b200 import capi
    p = problems.poissong("p1", 40)
    ref = oracle_runner.run(p)
    a = (ref["csr_rowptr"], ref["csr_colind"], ref["csr_val"])
    x_ref, _, _ = orc.gmres_ilu(*a, ref["rhs"], restart=300, levels=2, rtol=1e-13)
    B = cuda_runner.build(ctx, p)
    try:
        cuda_runner.assemble(B)
        s = capi.Solver(ctx, method="cg", precond=precond, maxits=5000, rtol=1e-10)
        x, rep = B.matrix.solve(B.vector, s)
        s.close()
        assert rep["converged"] == 1, rep
        assert np.linalg.norm(x - x_ref) <= 1e-8 * np.linalg.norm(x_ref)
    finally:
        cuda_runner.close(B)


def test_cg_iteration_count_matches_oracle(ctx):
    from tfel_b200 import capi
    ref = oracle_runner.run(problems.poisson("p1", 64))
    a = (ref["csr_rowptr"], ref["csr_colind"], ref["csr_val"])
    x_ref, its_ref, _ = orc.cg_jacobi(*a, ref["rhs"], rtol=1e-8)
    s = capi.Solver(ctx, method="cg", precond="jacobi", maxits=5000, rtol=1e-8, check_every=1)
    s.set_operator(a)
    x, rep = s.solve(ref["rhs"])
    s.close()
    assert abs(int(rep["its"]) - its_ref) <= 2, (rep, its_ref)
    assert np.linalg.norm(x - x_ref) <= 1e-8 * np.linalg.norm(x_ref)


def test_ilu0_factors_match_oracle(ctx):
    """GMRES + ILU(0) on Stokes: same iteration count as the oracle's restatement of PETSc's
    left-preconditioned GMRES (the factors and triangular solves agree to rounding)."""
    from tfel_b200 import capi
    ref = oracle_runner.run(problems.stokes(8))
    a = (ref["csr_rowptr"], ref["csr_colind"], ref["csr_val"])
    x_ref, its_ref, _ = orc.gmres_ilu(*a, ref["rhs"], restart=200, levels=0, rtol=1e-8, maxits=2000)
    s = capi.Solver(ctx, method="gmres", precond="ilu0", maxits=2000, restart=200, rtol=1e-8)
    s.set_operator(a)
    x, rep = s.solve(ref["rhs"])
    s.close()
    assert rep["converged"] == 1, rep
    assert abs(int(rep["its"]) - its_ref) <= 2, (rep, its_ref)
    x13, _, _ = orc.gmres_ilu(*a, ref["rhs"], restart=400, levels=0, rtol=1e-13, maxits=4000)
    assert np.linalg.norm(x - x13) <= 1e-6 * np.linalg.norm(x13)


@pytest.mark.parametrize("problem,kw", [
    (problems.stokes(10), dict(method="gmres", precond="ilu0", restart=300)),
    # ILU(0) of the indefinite Navier-Stokes matrix is so ill-conditioned that a 1e-11 preconditioned residual
    # leaves a 1e-7 error (the oracle's ILU(0)-GMRES behaves the same): full GMRES with Jacobi scaling instead
    (problems.navier_stokes(8), dict(method="gmres", precond="jacobi", restart=900)),
    # the solver every reference program configures (solver.hpp:139-163: GMRES(1000) + ILU(ilufill = 2))
    (problems.stokes(10), dict(method="gmres", precond="ilu", ilufill=2, restart=1000)),
    (problems.tet("p2", 4), dict(method="cg", precond="block_jacobi", block_size=4)),
    (problems.tet("p1", 6), dict(method="gmres", precond="jacobi", restart=100)),
    (problems.laplace2(12), dict(method="gmres", precond="none", restart=50)),
], ids=["stokes", "ns", "stokes_gmres_ilu2", "tet_p2", "tet_p1", "laplace2"])
def test_solution_parity(ctx, problem, kw):
    from tfel_b200 import capi
    ref = oracle_runner.run(problem)
    a = (ref["csr_rowptr"], ref["csr_colind"], ref["csr_val"])
    # reference solution: sparse direct solve of the oracle's system (the oracle's ILU(k)-GMRES stagnates on the
    # indefinite Navier-Stokes matrix)
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    x_ref = spl.spsolve(sp.csr_matrix((a[2], a[1], a[0])).tocsc(), ref["rhs"])
    B = cuda_runner.build(ctx, problem)
    try:
        cuda_runner.assemble(B)
        s = capi.Solver(ctx, maxits=20000, rtol=1e-11, **kw)
        x, rep = B.matrix.solve(B.vector, s)
        s.close()
        assert rep["converged"] == 1, rep
        assert np.linalg.norm(x - x_ref) <= 1e-8 * np.linalg.norm(x_ref), rep
    finally:
        cuda_runner.close(B)


def test_navier_stokes_with_the_configured_solver(ctx):
    """configs[4] with the solver the reference program configures (test/navier_stokes_2d_p2_p1.cpp:246-254: GMRES(1000) +
    ILU(2), left preconditioned, rtol 1e-8) on the parity fixture's matrix (synthetic, large previous iterates: strongly
    indefinite). What is checked is that the device solver BEHAVES like the restated reference solver: same iteration count
    (+-2; measured 43 = 43) and the same iterate (measured 4e-4 relative; the incomplete factors of this matrix are nearly
    singular, so the preconditioned residual both solvers stop on says little about the error -- either iterate is O(1)
    away from the direct solve, which is why `test_solution_parity[ns]` uses Jacobi). The Newton systems of the real
    program are benign: examples/navier_stokes_2d_p2_p1.cpp converges quadratically with this solver
    (test_cpp_surface.py, profiles/r02_ns_n128_100steps.log)."""
    from tfel_b200 import capi
    problem = problems.navier_stokes(8)
    ref = oracle_runner.run(problem)
    a = (ref["csr_rowptr"], ref["csr_colind"], ref["csr_val"])
    x_orc, its_orc, _ = orc.gmres_ilu(*a, ref["rhs"], restart=1000, levels=2, rtol=1e-8, maxits=2000)
    B = cuda_runner.build(ctx, problem)
    try:
        cuda_runner.assemble(B)
        s = capi.Solver(ctx, method="gmres", precond="ilu", ilufill=2, restart=1000, maxits=2000, rtol=1e-8)
        x, rep = B.matrix.solve(B.vector, s)
        s.close()
        assert rep["converged"] == 1 and rep["ilu_levels"] == 2, rep
        assert abs(int(rep["its"]) - its_orc) <= 2, (rep, its_orc)
        assert np.linalg.norm(x - x_orc) <= 1e-2 * np.linalg.norm(x_orc)
    finally:
        cuda_runner.close(B)


def test_tet_p2_known_answer(ctx):
    """u = |x|^2 is reproduced exactly by the new tetrahedron P2 element (test/tetrahedron_cube.cpp problem)."""
    from tfel_b200 import capi
    p = problems.tet("p2", 4)
    B = cuda_runner.build(ctx, p)
    try:
        cuda_runner.assemble(B)
        s = capi.Solver(ctx, method="cg", precond="jacobi", maxits=5000, rtol=1e-13)
        x, rep = B.matrix.solve(B.vector, s)
        s.close()
        xyz = B.spaces[0].dof_coordinates()
        assert np.abs(x - (xyz ** 2).sum(1)).max() < 1e-9
    finally:
        cuda_runner.close(B)


def test_rank0_integrals(ctx):
    """form.hpp:87-125 known answers: int x_2 = 1/2 on the cube (test/tetrahedron_cube.cpp:56-58);
    int (sin x + sin y) = 4 pi on [0, pi]^2 (test/quadrature_2d.cpp:10-18)."""
    from tfel_b200 import capi
    m = capi.Mesh.cube(ctx, 1, 1, 1, 3, 3, 3)
    v = capi.integrate(m, "qfSym35pTet", [dict(factors=[("x", problems.x1)])])
    m.close()
    assert abs(v - 0.5) < 1e-13
    m = capi.Mesh.square(ctx, np.pi, np.pi, 40, 40)
    v = capi.integrate(m, "qf5pT", [dict(factors=[("x", problems.X.sin(problems.x0) + problems.X.sin(problems.x1))])])
    m.close()
    assert abs(v - 4 * np.pi) < 1e-9


def test_errors_are_reported(ctx):
    from tfel_b200 import capi
    m = capi.Mesh.square(ctx, 1, 1, 3, 3)
    with pytest.raises(capi.TfelCudaError):
        capi.Space(m, 7)
    sp = capi.Space(m, "p1")
    a = capi.Matrix(sp)
    with pytest.raises(capi.TfelCudaError):
        a.assemble("qfSym4pTet", [dict(test=(0, 0), trial=(0, 0))])   # tetrahedron rule on a triangle mesh
    with pytest.raises(capi.TfelCudaError):
        a.assemble("qf5pT", [dict(test=(1, 0), trial=(0, 0))])        # component out of range
    s = capi.Solver(ctx)
    with pytest.raises(capi.TfelCudaError):
        s.solve(np.zeros(4))                                            # no operator
    s.close(); a.close(); sp.close(); m.close()


def test_full_size_structure_properties(ctx):
    """Size-independent properties at a size the oracle does not run in seconds (n = 1024):
    dofs (n+1)^2, nnz 7 (n-1)^2 + 4n, 4n Dirichlet dofs, row sums of the stiffness matrix vanish on
    interior rows whose neighbours are all interior, symmetry of the interior block via x.Ay = y.Ax."""
    from tfel_b200 import capi
    n = 1024
    p = problems.poisson("p1", n)
    B = cuda_runner.build(ctx, p)
    try:
        cuda_runner.assemble(B)
        assert B.spaces[0].n_dof == (n + 1) ** 2
        assert B.matrix.nnz == 7 * (n - 1) ** 2 + 4 * n
        dofs, _ = B.spaces[0].dirichlet()
        assert dofs.shape[0] == 4 * n
        s = capi.Solver(ctx)
        s.set_operator(B.matrix)
        ones = np.ones(B.matrix.n_rows)
        y = s.spmv(ones)
        isd = np.zeros(B.matrix.n_rows, dtype=bool); isd[dofs] = True
        assert np.all(y[isd] == 1.0)
        assert np.abs(y[~isd]).max() < 1e-11
        rs = np.random.RandomState(0)
        u = rs.standard_normal(B.matrix.n_rows); w = rs.standard_normal(B.matrix.n_rows)
        u[isd] = 0; w[isd] = 0
        au = s.spmv(u); aw = s.spmv(w)
        assert abs(np.dot(w[~isd], au[~isd]) - np.dot(u[~isd], aw[~isd])) <= 1e-10 * abs(np.dot(w[~isd], au[~isd]))
        s.close()
    finally:
        cuda_runner.close(B)


@pytest.mark.parametrize("problem", [problems.poissong("p2", 14), problems.stokes(9), problems.tet("p1b", 4)],
                         ids=lambda p: p.name)
def test_owned_row_segments_equal_rows_of_the_full_form(ctx, problem):
    """The row-distributed form restricted to any rank's segments holds exactly those rows of the
    reference CSR, bit for bit (no second GPU needed for this part of the multi-GPU path)."""
    from tfel_b200 import capi, dist as tdist
    B = cuda_runner.build(ctx, problem)
    try:
        cuda_runner.assemble(B)
        rp, ci, val = B.matrix.csr()
        b_full = B.vector.download()
        offs = [s.kind_offsets() for s in B.spaces]
        for P, rank in ((3, 0), (3, 1), (3, 2), (8, 5)):
            segs = tdist.strip_segments(offs, P, rank)
            A = capi.Matrix(B.spaces, B.spaces, rows=segs)
            b = capi.Vector(B.spaces, rows=segs)
            A.assemble(problem.quad, cuda_runner.lower(problem.bilinear), B.fields)
            b.assemble(problem.quad, cuda_runner.lower(problem.linear), B.fields)
            lrp, lci, lval = A.csr()
            rows = np.concatenate([np.arange(s[0], s[1]) for s in segs])
            sel = np.concatenate([np.arange(rp[g], rp[g + 1]) for g in rows])
            assert np.array_equal(np.diff(lrp), np.diff(rp)[rows])
            assert np.array_equal(lci, ci[sel])
            assert np.array_equal(lval, val[sel])
            assert np.array_equal(b.download(), b_full[rows])
            assert np.array_equal(A.make_rhs(b), B.matrix.make_rhs(B.vector)[rows])
            A.close(); b.close()
    finally:
        cuda_runner.close(B)


@pytest.mark.parametrize("name,problem", [("poisson_p1_n1024", problems.poisson("p1", 1024)),
                                          ("poisson_p1_n2048", problems.poisson("p1", 2048)),
                                          ("poisson_p1_n4096", problems.poisson("p1", 4096)),
                                          ("stokes_n256", problems.stokes(256)),
                                          ("tet_p1_n48", problems.tet("p1", 48))], ids=lambda v: v if isinstance(v, str) else "")
def test_full_size_fingerprints_match_the_reference(ctx, name, problem, golden_dir):
    """BASELINE.json sizes (configs[1] at n = 4096 included): DOF map, g2l, Dirichlet set, CSR rowptr / colind hash to
    the values the UNMODIFIED reference printed (tests/golden/refhash_*.json, oracle/gen_refhash.py); matrix, linear
    form and right-hand side agree in sum, sum of squares and sum of absolute values to 1e-11."""
    import json
    import refhash
    ref = json.load(open(os.path.join(golden_dir, "refhash_%s.json" % name)))
    got = cuda_runner.run(ctx, problem)
    try:
        assert got["csr_colind"].shape[0] == ref["nnz"]
        if not problem.composite and problem.dim == 2:   # the hypotenuse couplings the reference stores as exact zeros
            assert int((np.abs(got["csr_val"]) < 1e-13).sum()) == ref["n_exact_zero"]
        n = refhash.check(got, ref, None)
        assert n >= 8, n
    finally:
        cuda_runner.close(got["_built"])


@pytest.mark.parametrize("name,dim,n", [("mesh_square_n1024", 2, 1024), ("mesh_square_n4096", 2, 4096),
                                        ("mesh_cube_n48", 3, 48), ("mesh_cube_n142", 3, 142)])
def test_full_size_mesh_fingerprints_match_the_reference(ctx, name, dim, n, golden_dir):
    """a1-a4 on the device at the sizes of configs[1] (n = 4096 square) and configs[2] (n = 142 cube): vertices, cells, |K|,
    J^{-T}, face neighbours and the boundary submesh hash to what the UNMODIFIED reference printed
    (tests/golden/refhash_mesh_*.json, oracle/gen_refhash.py)."""
    import json
    import refhash
    from tfel_b200 import capi
    ref = json.load(open(os.path.join(golden_dir, "refhash_%s.json" % name)))
    m = capi.Mesh.square(ctx, 1.0, 1.0, n, n) if dim == 2 else capi.Mesh.cube(ctx, 1.0, 1.0, 1.0, n, n, n)
    try:
        got = {}
        got["vertices"], got["cells"] = m.download()
        got["cell_volume"], got["jmt"] = m.geometry()
        got["boundary_cells"], got["boundary_parent_cell"], got["boundary_parent_subdomain"] = m.boundary()
        got["cell_neighbours"] = m.neighbours()
        assert refhash.check(got, ref, None) == 8
    finally:
        m.close()


def test_quadrature_points_and_kind_offsets(ctx):
    """tfelcu_mesh_quadrature_points (cell.hpp:456-468) against the oracle; entity-kind offsets of fes.hpp:33-72."""
    from tfel_b200 import capi
    m = capi.Mesh.square(ctx, 1.0, 1.0, 7, 7)
    v, c = m.download()
    assert np.abs(m.quadrature_points("qf5pT") - orc.quadrature_points(v, c, "qf5pT")).max() <= 1e-15
    p2 = capi.Space(m, "p2"); p1 = capi.Space(m, "p1"); p1b = capi.Space(m, "p1b")
    assert p1.kind_offsets() == [0, 64, 64, 64, 64]
    assert p2.kind_offsets() == [0, 64, 64 + 161, 64 + 161, 64 + 161][:2] + [225, 225, 225]
    assert p1b.kind_offsets() == [0, 64, 64, 64 + 98, 64 + 98]
    for s in (p1, p2, p1b):
        s.close()
    m.close()
    m = capi.Mesh.cube(ctx, 1.0, 1.0, 1.0, 3, 3, 3)
    v, c = m.download()
    a = m.quadrature_points("qfSym10pTet"); b = orc.quadrature_points(v, c, "qfSym10pTet")
    # the order of the points inside a symmetry orbit is ours: compare per cell as sets
    assert np.abs(np.sort(a.reshape(a.shape[0], -1), axis=1) - np.sort(b.reshape(b.shape[0], -1), axis=1)).max() <= 1e-15
    m.close()


def test_vertex_and_cell_values(ctx):
    """to_mesh_vertex_data / to_mesh_cell_data on the device (fes.hpp:464-504): the vertex dofs per vertex, the discrete
    function at the barycentre of every cell -- against numpy on the downloaded maps and the oracle's basis tables."""
    rng = np.random.RandomState(11)
    for p in (problems.poisson("p1", 5), problems.poisson("p2", 7), problems.tet("p1b", 3), problems.tet("p2", 3)):
        B = cuda_runner.build(ctx, p)
        try:
            sp = B.spaces[0]
            dm = sp.dof_map().astype(np.int64)
            _, cells = B.mesh.download()
            dim = B.mesh.dim
            coef = rng.standard_normal(sp.n_dof)
            ref_v = np.zeros(B.mesh.n_vertices)
            ref_v[cells.astype(np.int64)] = coef[dm[:, :dim + 1]]
            assert np.array_equal(sp.vertex_values(coef), ref_v)
            bc = np.full((1, dim), 1.0 / (dim + 1))
            phi = orc.fe_tabulate(dim, orc.FE_BY_NAME[p.comps[0]], bc)[0, :, 0]
            ref_c = np.zeros(B.mesh.n_cells)
            for i in range(dm.shape[1]):
                ref_c = ref_c + coef[dm[:, i]] * phi[i]
            assert np.abs(sp.cell_values(coef) - ref_c).max() <= 1e-14 * np.abs(ref_c).max()
        finally:
            cuda_runner.close(B)


def test_mesh_with_unused_vertices_and_unsorted_cells(ctx):
    """Edge cases of the mesh / space constructors: vertices no cell uses are skipped by the numbering
    (get_vertex_list, cell.hpp:393-405: dofs are ranks among the USED vertices); cells whose first nv - 1 ids are not
    ascending are sorted like mesh.hpp:294-298 does; a cell that is still not ascending is refused
    (subdomain.hpp:11-13)."""
    from tfel_b200 import capi
    v, c = orc.gen_square_mesh(1.0, 1.0, 5, 5)
    keep = np.ones(c.shape[0], dtype=bool)
    keep[[0, 1, 2, 3, 10, 11]] = False          # vertex 0 (and 1) lose all / some of their cells
    c2 = np.ascontiguousarray(c[keep])
    for fe_name, fe in (("p1", orc.FE_P1), ("p2", orc.FE_P2), ("p1b", orc.FE_P1_BUBBLE)):
        m = capi.Mesh.create(ctx, v, c2)
        sp = capi.Space(m, fe_name)
        dof_map, N, g2l = orc.fes_build(2, fe, c2)
        assert sp.n_dof == N
        assert_int_equal(sp.dof_map(), dof_map, fe_name)
        assert_int_equal(sp.g2l(), g2l, fe_name)
        bel, bid, bsd = orc.boundary_submesh(2, c2)
        gel, gid, gsd = m.boundary()
        assert_int_equal(gel, bel); assert_int_equal(gid, bid); assert_int_equal(gsd, bsd)
        assert_int_equal(sp.boundary_dofs(), orc.dirichlet_dofs(2, fe, c2, bel))
        sp.close(); m.close()
    # first two ids swapped: sorted like the reference; result equals the sorted mesh
    c3 = c.copy(); c3[7, [0, 1]] = c3[7, [1, 0]]
    m = capi.Mesh.create(ctx, v, c3)
    assert np.array_equal(m.download()[1], c)
    m.close()
    # last id out of order: the reference only sorts the first nv - 1 ids and then throws
    c4 = c.copy(); c4[7, [1, 2]] = c4[7, [2, 1]]
    with pytest.raises(capi.TfelCudaError, match="ordered"):
        capi.Mesh.create(ctx, v, c4)


def test_empty_and_tiny_inputs(ctx):
    """One-cell meshes, a form with no terms, an integrand that is identically zero, a system of Dirichlet rows only."""
    from tfel_b200 import capi
    v = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]); c = np.array([[0, 1, 2]], dtype=np.uint32)
    m = capi.Mesh.create(ctx, v, c)
    sp = capi.Space(m, "p2")
    assert sp.n_dof == 6
    sp.add_dirichlet_boundary(1.5)                     # every dof of a single triangle is on the boundary
    a = capi.Matrix(sp); b = capi.Vector(sp)
    a.assemble("qf5pT", [dict(test=(0, 1), trial=(0, 1))])
    a.assemble("qf5pT", [])                            # no terms: nothing happens
    b.assemble("qf5pT", [dict(test=(0, 0), c=0.0)])
    rp, ci, val = a.csr()
    assert np.array_equal(rp, np.arange(7)) and np.array_equal(ci, np.arange(6)) and np.all(val == 1.0)
    s = capi.Solver(ctx, method="cg", precond="jacobi")
    x, rep = a.solve(b, s)
    assert rep["converged"] == 1 and rep["its"] == 0 and np.all(x == 1.5)
    s2 = capi.Solver(ctx, method="gmres", precond="ilu0", restart=5, maxits=10)
    x, rep = a.solve(b, s2)
    assert rep["converged"] == 1 and np.allclose(x, 1.5)
    assert capi.integrate(m, "qf5pT", [dict(c=2.0)]) == pytest.approx(1.0)   # 2 |K| = 1
    s.close(); s2.close(); a.close(); b.close(); sp.close(); m.close()


def test_beyond_2_30_stored_entries(ctx):
    """configs[3] at its full size (Stokes P2-P2-P1, n = 2048: 8 388 608 cells, 37 769 219 dofs, 1 119 416 369 stored
    entries, i.e. CSR positions beyond 2^30 where `lo + hi` of an int32 binary search overflows -- the bug this size
    exposed): structure, assembly of A and b, and the two SpMV kernels against each other on the assembled operator.
    The same steps as `tools/stokes_probe.py --assembly-only --spmv-check 2048` (about 3 s on a B200)."""
    import ctypes as C
    from tfel_b200 import capi
    n = 2048
    try:
        B = cuda_runner.build(ctx, problems.stokes(n))
    except capi.TfelCudaError as e:   # ~70 GB of device memory during the pattern build
        if "memory" in str(e):
            pytest.skip("not enough device memory for the full-size Stokes pattern")
        raise
    try:
        assert B.mesh.n_cells == 2 * n * n and B.matrix.n_rows == 2 * (2 * n + 1) ** 2 + (n + 1) ** 2
        assert B.matrix.nnz == 1119416369 and B.matrix.nnz > 2 ** 30
        cuda_runner.assemble(B)
        rowptr = np.zeros(B.matrix.n_rows + 1, dtype=np.int32)
        assert capi.lib().tfelcu_matrix_download_csr(B.matrix.h, rowptr.ctypes.data_as(C.c_void_p), None, None) == 0
        assert rowptr[0] == 0 and rowptr[-1] == B.matrix.nnz and np.all(np.diff(rowptr) >= 1)
        x = np.random.RandomState(3).standard_normal(B.matrix.n_rows)
        ys = []
        for kern in ("vector", "stream"):
            s = capi.Solver(ctx, method="gmres", precond="none", spmv=kern)
            s.set_operator(B.matrix)
            ys.append(s.spmv(x))
            s.close()
        assert np.all(np.isfinite(ys[0])) and np.abs(ys[0]).max() > 0.0
        assert np.abs(ys[0] - ys[1]).max() <= 1e-13 * np.abs(ys[0]).max()
    finally:
        cuda_runner.close(B)
