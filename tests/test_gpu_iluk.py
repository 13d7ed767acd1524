"""GPU: ILU(ilufill) in natural ordering -- PCILU + PCFactorSetLevels(pc, ilufill) of the reference
(src/core/solver.hpp:162-163; every reference program passes ilufill = 2) -- against the oracle's sequential
row-merge restatement: pattern bit-exact, factors to rounding, GMRES iteration counts within +-2, solutions to 1e-8
against a sparse direct solve of the same system."""
import numpy as np
import pytest

import cuda_runner
import oracle as orc
import oracle_runner
import problems

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from tfel_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _system(problem):
    ref = oracle_runner.run(problem)
    return (ref["csr_rowptr"], ref["csr_colind"], ref["csr_val"]), ref["rhs"]


def _direct(a, rhs):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    return spl.spsolve(sp.csr_matrix((a[2], a[1], a[0])).tocsc(), rhs)


@pytest.mark.parametrize("levels", [0, 1, 2, 3])
@pytest.mark.parametrize("problem", [problems.poisson("p1", 12), problems.poisson("p2", 7), problems.tet("p1", 4),
                                     problems.stokes(8), problems.navier_stokes(6)],
                         ids=lambda p: "%s_n%d" % (p.name, p.mesh[1]))
def test_iluk_pattern_and_factors_match_oracle(ctx, problem, levels):
    """symbolic phase (one fill-path search per row) == sequential row merge with the sum rule, bit for bit; numeric
    phase (row-parallel IKJ behind per-row flags) == the sequential elimination."""
    from tfel_b200 import capi
    a, _ = _system(problem)
    want_rp, want_ci, want_lu = orc.iluk(*a, levels)
    s = capi.Solver(ctx, method="gmres", precond="ilu", ilufill=levels)
    s.set_operator(a)
    rp, ci, lu = s.ilu_factors()
    s.close()
    assert np.array_equal(rp, want_rp.astype(np.int64))
    assert np.array_equal(ci, want_ci)
    finite = np.isfinite(want_lu)
    assert np.array_equal(np.isfinite(lu), finite)
    scale = np.abs(want_lu[finite]).max()
    assert np.abs(lu[finite] - want_lu[finite]).max() <= 1e-12 * scale


def test_iluk_random_unsymmetric_pattern(ctx):
    from tfel_b200 import capi
    rng = np.random.RandomState(7)
    n = 400
    mask = rng.rand(n, n) < 0.012
    np.fill_diagonal(mask, True)
    rp = np.zeros(n + 1, dtype=np.int32); rp[1:] = np.cumsum(mask.sum(1))
    ci = np.nonzero(mask)[1].astype(np.int32)
    val = rng.standard_normal(ci.shape[0])
    val[np.flatnonzero(ci == np.repeat(np.arange(n), np.diff(rp)))] += 8.0    # diagonally dominant: stable factors
    for levels in (0, 1, 2, 4):
        want_rp, want_ci, want_lu = orc.iluk(rp, ci, val, levels)
        s = capi.Solver(ctx, method="gmres", precond="ilu", ilufill=levels)
        s.set_operator((rp, ci, val))
        got_rp, got_ci, got_lu = s.ilu_factors()
        s.close()
        assert np.array_equal(got_rp, want_rp.astype(np.int64)), levels
        assert np.array_equal(got_ci, want_ci), levels
        assert np.abs(got_lu - want_lu).max() <= 1e-12 * np.abs(want_lu).max(), levels


@pytest.mark.parametrize("n", [32, 64, 128])
def test_stokes_gmres_ilu2_iterations_and_solution(ctx, n):
    """configs[3] shape with the reference's own solver settings (ilufill 2, restart 1000): iteration count of the
    oracle's restatement of PETSc's left-preconditioned GMRES within +-2, solution within 1e-8 of the direct solve."""
    from tfel_b200 import capi
    p = problems.stokes(n)
    a, rhs = _system(p)
    _, its_ref, _ = orc.gmres_ilu(*a, rhs, restart=1000, levels=2, rtol=1e-8, maxits=2000)
    B = cuda_runner.build(ctx, p)
    try:
        cuda_runner.assemble(B)
        s = capi.Solver(ctx, method="gmres", precond="ilu", ilufill=2, maxits=2000, restart=1000, rtol=1e-8)
        x, rep = B.matrix.solve(B.vector, s)
        assert rep["converged"] == 1, rep
        assert rep["ilu_levels"] == 2 and rep["ilu_nnz"] > 3 * B.matrix.nnz, rep
        assert abs(int(rep["its"]) - its_ref) <= 2, (rep, its_ref)
        s.close()
        s = capi.Solver(ctx, method="gmres", precond="ilu", ilufill=2, maxits=4000, restart=1000, rtol=1e-12)
        x, rep = B.matrix.solve(B.vector, s)
        s.close()
        assert rep["converged"] == 1, rep
        x_ref = _direct(a, rhs)
        assert np.linalg.norm(x - x_ref) <= 1e-8 * np.linalg.norm(x_ref), rep
    finally:
        cuda_runner.close(B)


def test_cg_ilu_levels_on_poisson(ctx):
    """SPD case: CG preconditioned by ILU(k); more fill, fewer iterations, same solution."""
    from tfel_b200 import capi
    p = problems.poissong("p1", 48)
    a, rhs = _system(p)
    x_ref = _direct(a, rhs)
    its = []
    for levels in (0, 1, 2):
        s = capi.Solver(ctx, method="cg", precond="ilu", ilufill=levels, maxits=3000, rtol=1e-11)
        s.set_operator(a)
        x, rep = s.solve(rhs)
        s.close()
        assert rep["converged"] == 1, rep
        assert np.linalg.norm(x - x_ref) <= 1e-8 * np.linalg.norm(x_ref), (levels, rep)
        its.append(int(rep["its"]))
    assert its[2] <= its[1] <= its[0], its


def test_values_only_refresh_reuses_the_symbolic_phase(ctx):
    """Newton / time loops (test/navier_stokes_2d_p2_p1.cpp:292-438): a new form on the same spaces and a new solver with
    the same parameters cost no structure build -- pattern, slot map, SpMV plan and ILU symbolic phase come from the
    parked objects; the factors follow the new values."""
    from tfel_b200 import capi
    p = problems.stokes(16)
    a, rhs = _system(p)
    B = cuda_runner.build(ctx, p)
    try:
        cuda_runner.assemble(B)
        s = capi.Solver(ctx, method="gmres", precond="ilu", ilufill=2, maxits=500, restart=500, rtol=1e-10)
        x1, rep1 = B.matrix.solve(B.vector, s)
        assert rep1["converged"] == 1 and rep1["t_symbolic_s"] > 0.0, rep1
        # same form, values doubled (a += the same integrand): the pattern is unchanged
        B.matrix.assemble(p.quad, cuda_runner.lower(p.bilinear))
        x2, rep2 = B.matrix.solve(B.vector, s)
        assert rep2["converged"] == 1 and rep2["t_symbolic_s"] == 0.0, rep2
        rp, ci, v2 = B.matrix.csr()
        want = orc.iluk(rp, ci, v2, 2)
        s.set_operator(B.matrix)
        got = s.ilu_factors()
        assert np.array_equal(got[1], want[1])
        assert np.abs(got[2] - want[2]).max() <= 1e-12 * np.abs(want[2]).max()
        s.close()
        # a new solver with the same parameters takes the parked one
        s2 = capi.Solver(ctx, method="gmres", precond="ilu", ilufill=2, maxits=500, restart=500, rtol=1e-10)
        x3, rep3 = B.matrix.solve(B.vector, s2)
        s2.close()
        assert rep3["converged"] == 1 and rep3["t_symbolic_s"] == 0.0, rep3
        assert np.array_equal(x2, x3)
    finally:
        cuda_runner.close(B)


def test_solver_keeps_its_own_copy_of_the_operator(ctx):
    """set_operator copies the matrix like the reference does (solver.cpp:71-96): clearing or destroying the form
    afterwards does not change what the solver solves; a solver that only borrowed the arrays refuses to run."""
    from tfel_b200 import capi
    p = problems.poisson("p1", 24)
    a, rhs = _system(p)
    B = cuda_runner.build(ctx, p)
    cuda_runner.assemble(B)
    s = capi.Solver(ctx, method="cg", precond="jacobi", maxits=2000, rtol=1e-10)
    s.set_operator(B.matrix)
    B.matrix.clear()
    B.matrix.csr()                     # performs the pending clear: the form's values are identity rows / zeros now
    x, rep = s.solve(rhs)
    assert rep["converged"] == 1
    x_ref = _direct(a, rhs)
    assert np.linalg.norm(x - x_ref) <= 1e-8 * np.linalg.norm(x_ref)
    cuda_runner.assemble(B)
    xs, _ = B.matrix.solve(B.vector, s)   # borrows the form's arrays
    with pytest.raises(capi.TfelCudaError):
        s.solve(rhs)
    cuda_runner.close(B)
    s.close()
