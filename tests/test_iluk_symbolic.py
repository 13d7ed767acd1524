"""Host logic for the next preconditioner step: the row-parallel (fill-path) formulation of symbolic ILU(k) in
tools/iluk_symbolic.py gives exactly the pattern of the sequential row-merge the oracle restates for PETSc's
PCFactorSetLevels (solver.hpp:162-163) -- on FE matrices whose patterns are NOT symmetric (identity Dirichlet rows)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import iluk_symbolic  # noqa: E402
import oracle as orc  # noqa: E402
import oracle_runner  # noqa: E402
import problems  # noqa: E402


@pytest.mark.parametrize("problem", [problems.poisson("p1", 7), problems.poisson("p2", 4), problems.tet("p1", 3), problems.stokes(4)],
                         ids=lambda p: "%s_n%d" % (p.name, p.mesh[1]))
@pytest.mark.parametrize("levels", [0, 1, 2, 3])
def test_fill_path_search_equals_sequential_row_merge(problem, levels):
    ref = oracle_runner.run(problem)
    rp, ci = ref["csr_rowptr"], ref["csr_colind"]
    want_rp, want_ci = orc.iluk_pattern(rp, ci, levels)
    got_rp, got_ci = iluk_symbolic.iluk_pattern(rp, ci, levels)
    assert np.array_equal(got_rp, want_rp)
    assert np.array_equal(got_ci, want_ci)
    if levels == 0:
        assert np.array_equal(got_rp, rp) and np.array_equal(got_ci, ci)


def test_random_unsymmetric_patterns():
    rng = np.random.RandomState(4)
    for n, density in ((30, 0.08), (60, 0.05)):
        mask = rng.rand(n, n) < density
        np.fill_diagonal(mask, True)
        rp = np.zeros(n + 1, dtype=np.int32); rp[1:] = np.cumsum(mask.sum(1))
        ci = np.nonzero(mask)[1].astype(np.int32)
        for levels in (0, 1, 2, 4):
            want = orc.iluk_pattern(rp, ci, levels)
            got = iluk_symbolic.iluk_pattern(rp, ci, levels)
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (n, levels)
