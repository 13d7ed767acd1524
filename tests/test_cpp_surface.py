"""The C++14 face (include/tfel/tfel.hpp): the reference's own programs compile against it unchanged except for the
solver plugin (CPU: compile + link only; GPU: run and compare with the oracle)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "examples", "bin")


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tfel_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "examples")])


def test_reference_programs_compile_against_the_cuda_headers():
    build()
    for name in ("poisson", "stokes_2d_p2_p1", "navier_stokes_2d_p2_p1", "plugin_and_forms", "expression_algebra", "device_functor", "poisson_device", "projectors",
                 "export_ensight", "export_format_host", "laplacian_transient"):
        assert os.path.exists(os.path.join(BIN, name))


REFERENCE = "/root/reference"
# programs of the reference that compile UNCHANGED against include/tfel/tfel.hpp (with TFEL_CUDA_AS_PETSC their
# `solver::petsc::gmres_ilu s(p);` line runs the device GMRES + ILU(0))
UNCHANGED_REFERENCE_PROGRAMS = ("laplacian_transient", "tetrahedron_cube", "stokes_2d_p2_p1", "navier_stokes_2d_p2_p1",
                                "fe_derivative_form", "quadrature_2d", "l2_p1_bubble_projection")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "test")), reason="needs the reference tree (build container only)")
@pytest.mark.parametrize("name", UNCHANGED_REFERENCE_PROGRAMS)
def test_reference_test_programs_compile_unchanged(name, tmp_path):
    """Source compatibility of the drop-in surface, checked on the reference's own test programs: the file is copied
    (into a scratch directory, never into the repo) next to forwarding headers src/core/*.hpp -> <tfel/tfel.hpp>, and
    has to pass g++ -std=c++14 -fsyntax-only as it is."""
    import shutil
    core = tmp_path / "src" / "core"
    core.mkdir(parents=True)
    for h in os.listdir(os.path.join(REFERENCE, "src", "core")):
        if h.endswith(".hpp"):
            (core / h).write_text("#include <tfel/tfel.hpp>\n")
    (tmp_path / "spikes").mkdir()
    (tmp_path / "spikes" / "timer.hpp").write_text(
        "#pragma once\n#include <chrono>\nclass timer { public: timer() : t0(std::chrono::steady_clock::now()) {}\n"
        "  double tic() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }\n"
        "private: std::chrono::steady_clock::time_point t0; };\n")
    (tmp_path / "test").mkdir()
    shutil.copy(os.path.join(REFERENCE, "test", name + ".cpp"), str(tmp_path / "test" / (name + ".cpp")))
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-w", "-DTFEL_CUDA_AS_PETSC", "-I", os.path.join(ROOT, "include"),
                        "-I", str(tmp_path), str(tmp_path / "test" / (name + ".cpp"))], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]


EXPORT_GOLDEN = os.path.join(ROOT, "tests", "golden", "export")   # written by the UNMODIFIED reference (oracle/ref_driver.cpp)


def _compare_export(tmp_path, program, expect_all):
    (tmp_path / "export").mkdir()
    r = subprocess.run([os.path.join(BIN, program), "export"], capture_output=True, text=True, cwd=str(tmp_path), timeout=300)
    assert r.returncode == 0 and program + ": ok" in r.stdout, r.stdout + r.stderr
    written = sorted(os.listdir(str(tmp_path / "export")))
    golden = sorted(os.listdir(EXPORT_GOLDEN))
    assert written and set(written) <= set(golden), (written, golden)
    if expect_all:
        assert written == golden
    for name in written:
        got = open(os.path.join(str(tmp_path), "export", name), "rb").read()
        ref = open(os.path.join(EXPORT_GOLDEN, name), "rb").read()
        assert got == ref, name


def test_exporters_write_the_reference_files_host_only(tmp_path):
    """exporter::ensight6 / ensight6_geometry / ensight6_transient / ascii (export.hpp) on a host mesh: byte-identical to
    the files the unmodified reference wrote for the same data (tests/golden/export)."""
    build()
    _compare_export(tmp_path, "export_format_host", expect_all=False)


@pytest.mark.gpu
def test_exporters_and_mesh_data_on_the_cuda_path(tmp_path):
    """to_mesh_vertex_data / to_mesh_cell_data (fes.hpp:464-504) gathered on the device + the exporters: every file of the
    reference's run (P1 / P2 triangles, tetrahedra, vector data, transient series), byte for byte."""
    build()
    _compare_export(tmp_path, "export_ensight", expect_all=True)


def test_expression_algebra_matches_reference_assertions():
    """test/expression.cpp:20-50 of the reference (the only assertions in its test tree) on the host-side lowering."""
    build()
    r = subprocess.run([os.path.join(BIN, "expression_algebra")], capture_output=True, text=True)
    assert r.returncode == 0 and "expression_algebra: ok" in r.stdout, r.stdout + r.stderr


def test_no_device_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    build()
    r = subprocess.run([os.path.join(BIN, "poisson"), "4"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_readme_poisson_program(tmp_path):
    import oracle as orc
    import oracle_runner
    import problems
    build()
    out = str(tmp_path / "u.bin")
    r = subprocess.run([os.path.join(BIN, "poisson"), "10", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "dofs 121 dirichlet 40" in r.stdout     # README case
    ref = oracle_runner.run(problems.poisson("p1", 10))
    x_ref, _, _ = orc.cg_jacobi(ref["csr_rowptr"], ref["csr_colind"], ref["csr_val"], ref["rhs"], rtol=1e-13)
    x = np.fromfile(out)
    assert np.linalg.norm(x - x_ref) <= 1e-8 * np.linalg.norm(x_ref)
    r = subprocess.run([os.path.join(BIN, "poisson"), "256"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert abs(float(r.stdout.split("max_u")[1]) - 0.0736713) < 2e-5


@pytest.mark.gpu
def test_reference_stokes_program(tmp_path):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    import oracle_runner
    import problems
    build()
    out = str(tmp_path / "x.bin")
    r = subprocess.run([os.path.join(BIN, "stokes_2d_p2_p1"), "10", out, "1e-12"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "converged 1" in r.stdout
    ref = oracle_runner.run(problems.stokes(10))
    x_ref = spl.spsolve(sp.csr_matrix((ref["csr_val"], ref["csr_colind"], ref["csr_rowptr"])).tocsc(), ref["rhs"])
    x = np.fromfile(out)
    assert np.linalg.norm(x - x_ref) <= 1e-8 * np.linalg.norm(x_ref)


@pytest.mark.gpu
def test_user_plugin_fields_and_integrals():
    build()
    r = subprocess.run([os.path.join(BIN, "plugin_and_forms")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "plugin_and_forms: ok" in r.stdout


@pytest.mark.gpu
def test_reference_navier_stokes_program():
    """test/navier_stokes_2d_p2_p1.cpp:236-448 on the CUDA path: Newton iterations converge (relative velocity step
    below 1e-8) within the reference's 10 steps for the first time steps; per-step reassembly with coefficient fields."""
    build()
    r = subprocess.run([os.path.join(BIN, "navier_stokes_2d_p2_p1"), "12", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr
    assert "warning" not in r.stdout
    steps = [line for line in r.stdout.splitlines() if line.startswith("newton step #")]
    assert len(steps) >= 4
    # quadratic convergence: the last Newton step of each time step is below the tolerance
    last = [float(line.split("(")[1].split(",")[0]) for line in steps]
    assert min(last) < 1e-8
    assert "summary:" in r.stdout
    # the configured solver is the reference's: GMRES(1000) + ILU(ilufill = 2); after the first Newton step the per-step
    # bilinear_form and solver find their structure parked (pattern, slot map, SpMV plan, ILU symbolic phase)
    import json
    js = [json.loads(line[6:]) for line in r.stdout.splitlines() if line.startswith("json: ")]
    assert len(js) == 1 and js[0]["linear_solves_not_converged"] == 0
    st = js[0]["structure_ms"]
    assert st["first_step"]["ilu_symbolic"] > 0.0 and st["later_steps_mean"]["ilu_symbolic"] == 0.0, st
    assert st["later_steps_mean"]["form_constructor"] < 0.5 * st["first_step"]["form_constructor"] + 0.5, st


@pytest.mark.gpu
def test_device_functor_coefficients():
    """include/tfel/tfel_device.cuh: coefficient functions as __device__ functors instantiated in the user's own
    (nvcc-compiled) translation unit give the same system as the host-callback path."""
    build()
    r = subprocess.run([os.path.join(BIN, "device_functor"), "40"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "device_functor: ok" in r.stdout


@pytest.mark.gpu
def test_poisson_with_a_device_source_makes_no_host_callbacks(tmp_path):
    """examples/poisson.cpp's Gaussian source as a __device__ functor evaluated by a kernel of the user's module at += time
    (TFELCU_FACTOR_FUNCTION): zero host evaluations of integrand leaves, and the linear form equals the oracle's with the
    same source tabulated by numpy (1e-12)."""
    import numpy as np
    import oracle_runner
    import problems
    build()
    n = 40
    out = str(tmp_path / "b.bin")
    r = subprocess.run([os.path.join(BIN, "poisson_device"), str(n), out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host evaluations of integrand leaves 0," in r.stdout
    b = np.fromfile(out)
    p = problems.poisson("p1", n)
    p.linear = [dict(test=(0, 0), c=1.0, factors=[("x", problems.gaussian_src)])]
    ref = oracle_runner.run(p)
    assert np.abs(b - ref["linear_form"]).max() <= 1e-12 * np.abs(ref["linear_form"]).max()


@pytest.mark.gpu
def test_reference_transient_laplacian_program(tmp_path):
    """test/laplacian_transient.cpp on the CUDA path: one bilinear form solved at every implicit Euler step, the right-hand
    side reassembled from the previous iterate, every step exported (ensight6_transient); the exact solution
    (1 - t) sin(pi x) sin(pi y) is reproduced up to the spatial error."""
    build()
    r = subprocess.run([os.path.join(BIN, "laplacian_transient"), "40", "10", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "laplacian_transient: ok" in r.stdout, r.stdout + r.stderr
    case = open(os.path.join(str(tmp_path), "solution.case")).read()
    assert "number of steps: 11" in case and "scalar per node: 1 c" in case
    assert os.path.exists(os.path.join(str(tmp_path), "solution.var.c.000010"))
    assert os.path.exists(os.path.join(str(tmp_path), "initial_condition.var.c_init"))


@pytest.mark.gpu
@pytest.mark.parametrize("n", [16, 40])
def test_projectors_and_element_algebra_against_the_oracle(tmp_path, n):
    """projector::lagrange / projector::l2 (projector.hpp:12-97) and the element algebra (fes.hpp:314-405) of
    examples/projectors.cpp against the oracle: nodal values at the oracle's dof coordinates (exact), mass-matrix solves
    of the oracle's own M and rhs (direct solve, 1e-6: the projector stops CG at rtol 1e-8), linear combinations to
    rounding; the program itself checks that the coefficients never left the device on the solve -> algebra -> +=
    -> solve chain."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    import oracle as orc
    build()
    r = subprocess.run([os.path.join(BIN, "projectors"), str(n)], capture_output=True, text=True, cwd=str(tmp_path), timeout=300)
    assert r.returncode == 0 and "projectors: ok" in r.stdout, r.stdout + r.stderr
    assert r.stdout.count("coefficient transfers on the device chain: 0") == 2
    data = np.fromfile(os.path.join(str(tmp_path), "projectors.bin"))
    vertices, cells = orc.gen_square_mesh(1.0, 1.0, n, n)
    quad = "qf5pT"
    xq = orc.quadrature_points(vertices, cells, quad)
    wave = lambda x: np.sin(3.0 * x[..., 0]) * np.cos(2.0 * x[..., 1]) + 0.25
    bump = lambda x: x[..., 0] * (1.0 - x[..., 0]) * x[..., 1] + 0.5 * x[..., 1] * x[..., 1]
    at = 0
    for fe_name in ("p1", "p2"):
        fe = orc.FE_BY_NAME[fe_name]
        dof_map, N, g2l = orc.fes_build(2, fe, cells)
        got = [data[at + i * N: at + (i + 1) * N] for i in range(7)]
        at += 7 * N
        a, b, pa, pb, w, pw, mix = got
        x = orc.dof_coordinates(2, fe, vertices, cells, g2l, np.arange(N, dtype=np.uint32))
        assert np.array_equal(a, wave(x)) and np.array_equal(b, bump(x))
        space = orc.Space(2, [(fe, dof_map, N)])
        rowptr, colind = orc.pattern(cells.shape[0], space, space, None)
        val = orc.assemble_bilinear(vertices, cells, space, space, quad, [dict(test=(0, 0), trial=(0, 0), c=1.0)], None, False, rowptr, colind)
        M = sp.csr_matrix((val, colind, rowptr)).tocsc()
        lu = spl.splu(M)

        def l2(table):
            return lu.solve(orc.assemble_linear(vertices, cells, space, quad, [dict(test=(0, 0), c=1.0, table=table)], False))

        def close(u, v, tol):
            return np.abs(u - v).max() <= tol * np.abs(v).max()
        assert close(pa, l2(wave(xq)), 1e-6)
        assert close(pb, l2(bump(xq)), 1e-6)
        assert close(w, 2.0 * pa - pb / 3.0 + pa, 1e-14)
        assert close(mix, 0.5 * a + b * 3.0 - a / 4.0, 1e-14)
        # the next form takes w and d_x pb as coefficient fields (device arrays of the elements above)
        table = (orc.field_at_quadrature(fe, vertices, cells, dof_map, w, 0, quad) +
                 orc.field_at_quadrature(fe, vertices, cells, dof_map, pb, 1, quad))
        assert close(pw, l2(table), 1e-6)
    assert at == data.shape[0]
