"""GPU: the two fast paths for constant-coefficient forms on P1 triangles -- the closed-form rows of p1rows.cu (what AUTO
picks: scatter=None) and the tiles of tile.cu (scatter="tile"), both with the linear form riding along in the same pass --
against the oracle, against the generic owner-computes rows (scatter="rows"), and against themselves (bitwise
reproducible)."""
import numpy as np
import pytest

import cuda_runner
import oracle as orc
import oracle_runner
import problems
from parity import assert_int_equal, assert_values_close, matrix_entries_close

pytestmark = pytest.mark.gpu

STIFF = [dict(test=(0, 1), trial=(0, 1)), dict(test=(0, 2), trial=(0, 2))]
MASS = [dict(test=(0, 0), trial=(0, 0), c=0.75)]
MIXED = STIFF + MASS + [dict(test=(0, 1), trial=(0, 0), c=0.3), dict(test=(0, 0), trial=(0, 2), c=-0.2),
                        dict(test=(0, 1), trial=(0, 2), c=0.1)]
SRC = [dict(test=(0, 0), c=1.0)]


@pytest.fixture(scope="module")
def ctx():
    from tfel_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


# None = what AUTO picks (p1rows.cu: vertex tables per block + the row kernel); the two environment knobs select its other
# kernels: the streamed one (TMA producer warp) and the one without vertex tables (meshes whose blocks refer to more than
# 512 vertices fall back to it)
FAST = ["tile", None, "auto+TFELCU_P1STREAM", "auto+TFELCU_NO_VTAB"]
FAST_IDS = ["tile", "auto_p1rows", "p1rows_streamed", "p1rows_offsets"]


def _pick(scatter, monkeypatch):
    """scatter argument of Matrix(...) for one FAST entry; sets the environment knob of the entry for this test"""
    if scatter is not None and scatter.startswith("auto+"):
        monkeypatch.setenv(scatter.split("+")[1], "1")
        monkeypatch.setenv("TFELCU_NO_POOL", "1")     # structures built under another knob must not be re-issued
        return None
    return scatter


@pytest.mark.parametrize("scatter", FAST, ids=FAST_IDS)
@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 57, 130, 257])
def test_tile_matches_oracle(ctx, n, scatter, monkeypatch):
    """configs[0] / [1] shape at sizes around the tile edges (16 x 16 vertices per tile): structure bit-exact, values and
    the fused right-hand side to 1e-12"""
    scatter = _pick(scatter, monkeypatch)
    p = problems.poisson("p1", n)
    ref = oracle_runner.run(p)
    got = cuda_runner.run(ctx, p, scatter=scatter)
    try:
        for key in ("csr_rowptr", "csr_colind"):
            assert_int_equal(got[key], ref[key], key)
        matrix_entries_close(ref["csr_rowptr"], got["csr_val"], ref["csr_val"], 1e-12)
        assert_values_close(got["linear_form"], ref["linear_form"], 1e-12, "linear_form")
        assert_values_close(got["rhs"], ref["rhs"], 1e-12, "rhs")
    finally:
        cuda_runner.close(got["_built"])


def _forms(ctx, n, scatter, jitter=0.0, dirichlet=True):
    from tfel_b200 import capi
    v, c = orc.gen_square_mesh(1.0, 1.0, n, n)
    if jitter:
        rs = np.random.RandomState(5)
        inner = (v[:, 0] > 1e-9) & (v[:, 0] < 1 - 1e-9) & (v[:, 1] > 1e-9) & (v[:, 1] < 1 - 1e-9)
        v = v.copy(); v[inner] += (rs.rand(int(inner.sum()), 2) - 0.5) * (jitter / n)
    m = capi.Mesh.create(ctx, v, c)
    sp = capi.Space(m, "p1")
    if dirichlet:
        sp.add_dirichlet_boundary(0.0)
    return m, sp, capi.Matrix(sp, sp, scatter=scatter), capi.Vector(sp)


ANISO = [dict(test=(0, 1), trial=(0, 1), c=2.0), dict(test=(0, 2), trial=(0, 2), c=0.5), dict(test=(0, 1), trial=(0, 2), c=0.25)]


@pytest.mark.parametrize("fast", FAST, ids=FAST_IDS)
@pytest.mark.parametrize("terms", [STIFF, MASS, STIFF + MASS, ANISO, MIXED],
                         ids=["stiffness", "mass", "stiffness_plus_mass", "anisotropic", "all_tensor_blocks"])
@pytest.mark.parametrize("dirichlet", [True, False])
def test_tile_equals_rows_on_a_distorted_mesh(ctx, terms, dirichlet, fast, monkeypatch):
    """general triangles (jittered vertices), every block of the constant-coefficient tensor (value x value, value x
    gradient, gradient x value, gradient x gradient): tile == owner-computes rows up to the order of the additions"""
    fast = _pick(fast, monkeypatch)
    out = {}
    for key, scatter in (("rows", "rows"), ("fast", fast)):
        m, sp, A, b = _forms(ctx, 75, scatter, jitter=0.6, dirichlet=dirichlet)
        A.assemble("qf5pT", terms)
        b.assemble("qf5pT", SRC)
        out[key] = (A.csr()[2], b.download())
        b.close(); A.close(); sp.close(); m.close()
    scale = np.abs(out["rows"][0]).max()
    assert np.abs(out["fast"][0] - out["rows"][0]).max() <= 2e-14 * scale
    assert np.abs(out["fast"][1] - out["rows"][1]).max() <= 2e-14 * np.abs(out["rows"][1]).max()


@pytest.mark.parametrize("scatter", FAST, ids=FAST_IDS)
def test_fast_paths_are_bitwise_reproducible(ctx, scatter, monkeypatch):
    scatter = _pick(scatter, monkeypatch)
    runs = []
    for _ in range(2):
        m, sp, A, b = _forms(ctx, 200, scatter, jitter=0.4)
        A.assemble("qf5pT", MIXED)
        b.assemble("qf5pT", SRC)
        runs.append((A.csr()[2].copy(), b.download().copy()))
        b.close(); A.close(); sp.close(); m.close()
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])


@pytest.mark.parametrize("scatter", FAST, ids=FAST_IDS)
def test_fused_linear_form_equals_separate_pass(ctx, scatter, monkeypatch):
    """b += ... right after a += ... rides along in the same pass; looking at the matrix in between launches the += on
    its own and the linear form takes its usual kernel: same vector up to the order of the additions"""
    scatter = _pick(scatter, monkeypatch)
    m, sp, A, b = _forms(ctx, 90, scatter, jitter=0.5)
    A.assemble("qf5pT", STIFF)
    b.assemble("qf5pT", [dict(test=(0, 0), c=2.5)])
    fused = b.download()
    v_fused = A.csr()[2]
    A.clear(); b.clear()
    A.assemble("qf5pT", STIFF)
    v_alone = A.csr()[2]                      # flushes the deferred +=
    b.assemble("qf5pT", [dict(test=(0, 0), c=2.5)])
    alone = b.download()
    assert np.array_equal(v_fused, v_alone)
    assert np.abs(fused - alone).max() <= 2e-14 * np.abs(alone).max()
    # a linear form with a derivative does not ride along but is still right
    A.clear(); b.clear()
    A.assemble("qf5pT", STIFF)
    b.assemble("qf5pT", [dict(test=(0, 1), c=1.0)])
    assert np.array_equal(A.csr()[2], v_alone)
    b.close(); A.close(); sp.close(); m.close()


@pytest.mark.parametrize("scatter", FAST, ids=FAST_IDS)
def test_tile_accumulate_twice_and_clear(ctx, scatter, monkeypatch):
    scatter = _pick(scatter, monkeypatch)
    m, sp, A, b = _forms(ctx, 40, scatter)
    A.assemble("qf5pT", STIFF)
    rp, ci, v1 = A.csr()
    A.assemble("qf5pT", STIFF)
    A.assemble("qf5pT", STIFF)               # two deferred += in a row: the first is launched by the second
    v3 = A.csr()[2]
    dofs, _ = sp.dirichlet()
    is_dir = np.zeros(v1.shape[0], dtype=bool); is_dir[rp[dofs]] = True
    assert np.array_equal(v3[~is_dir], 3.0 * v1[~is_dir]) and np.all(v3[is_dir] == 1.0)
    A.assemble("qf5pT", STIFF)
    A.clear()                                # a += x; clear(): x is gone
    v0 = A.csr()[2]
    assert np.all(v0[is_dir] == 1.0) and np.all(v0[~is_dir] == 0.0)
    b.close(); A.close(); sp.close(); m.close()


def test_variable_coefficients_fall_back(ctx):
    p = problems.poissong("p1", 33)
    ref = oracle_runner.run(p)
    got = cuda_runner.run(ctx, p)            # AUTO: the tile path refuses coefficient factors, rows take over
    try:
        matrix_entries_close(ref["csr_rowptr"], got["csr_val"], ref["csr_val"], 1e-12)
        assert_values_close(got["linear_form"], ref["linear_form"], 1e-12, "linear_form")
    finally:
        cuda_runner.close(got["_built"])
