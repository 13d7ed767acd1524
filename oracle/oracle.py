"""TEST INFRASTRUCTURE: ctypes front-end of the plain-C oracle (oracle/tfel_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may
import this module. The product package (tfel_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "tfel_ref")

FE_P1, FE_P1_BUBBLE, FE_P2 = 1, 2, 3
FE_BY_NAME = {"p1": FE_P1, "p1b": FE_P1_BUBBLE, "p2": FE_P2}


def build(with_reference=None):
    """Compile liboracle.so (and oracle/_ref/tfel_ref when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    if with_reference is None:
        with_reference = os.path.isdir("/root/reference/src/core")
    if with_reference:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class _Term(C.Structure):
    _fields_ = [("test_comp", C.c_int), ("test_d", C.c_int),
                ("trial_comp", C.c_int), ("trial_d", C.c_int),
                ("c", C.c_double), ("table", C.POINTER(C.c_double))]


class _Space(C.Structure):
    _fields_ = [("dim", C.c_int), ("n_comp", C.c_int), ("fe", C.POINTER(C.c_int)),
                ("dof_map", C.POINTER(C.POINTER(C.c_uint))),
                ("n_dof", C.POINTER(C.c_size_t))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(with_reference=False)
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_boundary_submesh.restype = C.c_size_t
        _lib.orc_fes_build.restype = C.c_size_t
        _lib.orc_fes_dirichlet_dofs.restype = C.c_size_t
        _lib.orc_pattern.restype = C.c_size_t
        _lib.orc_integrate.restype = C.c_double
        _lib.orc_free.argtypes = [C.c_void_p]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def fe_ndof(dim, fe):
    return lib().orc_fe_ndof(dim, fe)


def fe_nodes(dim, fe):
    out = np.zeros((fe_ndof(dim, fe), dim))
    lib().orc_fe_nodes(dim, fe, _p(out, C.c_double))
    return out


def fe_tabulate(dim, fe, pts):
    pts = _f64(pts).reshape(-1, dim)
    out = np.zeros((pts.shape[0], fe_ndof(dim, fe), dim + 1))
    lib().orc_fe_tabulate(dim, fe, pts.shape[0], _p(pts, C.c_double), _p(out, C.c_double))
    return out


def quadrature(name):
    dim = C.c_int(0)
    nq = lib().orc_quadrature(name.encode(), None, None, C.byref(dim))
    if nq < 0:
        raise KeyError(name)
    x = np.zeros((nq, dim.value)); w = np.zeros(nq)
    lib().orc_quadrature(name.encode(), _p(x, C.c_double), _p(w, C.c_double), C.byref(dim))
    return x, w


def gen_square_mesh(x1, x2, n1, n2):
    v = np.zeros(((n1 + 1) * (n2 + 1), 2)); c = np.zeros((2 * n1 * n2, 3), dtype=np.uint32)
    lib().orc_gen_square_mesh(C.c_double(x1), C.c_double(x2), C.c_uint(n1), C.c_uint(n2),
                              _p(v, C.c_double), _p(c, C.c_uint))
    return v, c


def gen_cube_mesh(x1, x2, x3, n1, n2, n3):
    v = np.zeros(((n1 + 1) * (n2 + 1) * (n3 + 1), 3)); c = np.zeros((5 * n1 * n2 * n3, 4), dtype=np.uint32)
    lib().orc_gen_cube_mesh(C.c_double(x1), C.c_double(x2), C.c_double(x3),
                            C.c_uint(n1), C.c_uint(n2), C.c_uint(n3),
                            _p(v, C.c_double), _p(c, C.c_uint))
    return v, c


def geometry(vertices, cells):
    vertices = _f64(vertices); cells = _u32(cells)
    dim = vertices.shape[1]; K = cells.shape[0]
    vol = np.zeros(K); jmt = np.zeros((K, dim, dim))
    lib().orc_geometry(dim, _p(vertices, C.c_double), _p(cells, C.c_uint), C.c_size_t(K),
                       _p(vol, C.c_double), _p(jmt, C.c_double))
    return vol, jmt


def _take(ptr, n, dtype):
    if n == 0:
        out = np.zeros(0, dtype=dtype)
    else:
        out = np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)
    lib().orc_free(C.cast(ptr, C.c_void_p))
    return out


def boundary_submesh(dim, cells):
    cells = _u32(cells)
    el = C.POINTER(C.c_uint)(); el_id = C.POINTER(C.c_uint)(); sd_id = C.POINTER(C.c_uint)()
    F = lib().orc_boundary_submesh(dim, _p(cells, C.c_uint), C.c_size_t(cells.shape[0]),
                                   C.byref(el), C.byref(el_id), C.byref(sd_id))
    return (_take(el, F * dim, np.uint32).reshape(F, dim), _take(el_id, F, np.uint32),
            _take(sd_id, F, np.uint32))


def cell_neighbours(dim, cells):
    cells = _u32(cells)
    nb = np.zeros((cells.shape[0], dim + 1), dtype=np.int32)
    lib().orc_cell_neighbours(dim, _p(cells, C.c_uint), C.c_size_t(cells.shape[0]), _p(nb, C.c_int))
    return nb


def fes_build(dim, fe, cells):
    cells = _u32(cells); K = cells.shape[0]; nd = fe_ndof(dim, fe)
    dof_map = np.zeros((K, nd), dtype=np.uint32)
    N = lib().orc_fes_build(dim, fe, _p(cells, C.c_uint), C.c_size_t(K), _p(dof_map, C.c_uint))
    g2l = np.zeros((N, 2), dtype=np.uint32)
    lib().orc_fes_g2l(_p(dof_map, C.c_uint), C.c_size_t(K), nd, C.c_size_t(N), _p(g2l, C.c_uint))
    return dof_map, int(N), g2l


def dirichlet_dofs(dim, fe, cells, bcells):
    cells = _u32(cells); bcells = _u32(bcells)
    if bcells.ndim == 1:
        bcells = bcells.reshape(-1, 1)
    out = C.POINTER(C.c_uint)()
    n = lib().orc_fes_dirichlet_dofs(dim, fe, _p(cells, C.c_uint), C.c_size_t(cells.shape[0]),
                                     _p(bcells, C.c_uint), C.c_size_t(bcells.shape[0]),
                                     bcells.shape[1], C.byref(out))
    return _take(out, n, np.uint32)


def dof_coordinates(dim, fe, vertices, cells, g2l, dofs):
    vertices = _f64(vertices); cells = _u32(cells); g2l = _u32(g2l); dofs = _u32(dofs)
    x = np.zeros((dofs.shape[0], dim))
    lib().orc_dof_coordinates(dim, fe, _p(vertices, C.c_double), _p(cells, C.c_uint),
                              _p(g2l, C.c_uint), _p(dofs, C.c_uint), C.c_size_t(dofs.shape[0]),
                              _p(x, C.c_double))
    return x


class Space:
    """A (composite) finite element space: list of (fe, dof_map, n_dof) on one mesh."""

    def __init__(self, dim, components):
        self.dim = dim
        self.fe = [c[0] for c in components]
        self.dof_maps = [_u32(c[1]) for c in components]
        self.n_dofs = [int(c[2]) for c in components]
        self.offsets = np.concatenate([[0], np.cumsum(self.n_dofs)]).astype(np.int64)
        self.total = int(self.offsets[-1])
        n = len(components)
        self._fe = (C.c_int * n)(*self.fe)
        self._maps = (C.POINTER(C.c_uint) * n)(*[_p(m, C.c_uint) for m in self.dof_maps])
        self._nd = (C.c_size_t * n)(*self.n_dofs)
        self.c = _Space(dim, n, self._fe, self._maps, self._nd)


def _terms(terms, keep):
    arr = (_Term * max(1, len(terms)))()
    for i, t in enumerate(terms):
        arr[i].test_comp, arr[i].test_d = t["test"]
        arr[i].trial_comp, arr[i].trial_d = t.get("trial", (-1, 0))
        arr[i].c = float(t.get("c", 1.0))
        tab = t.get("table")
        if tab is not None:
            tab = _f64(tab); keep.append(tab)
            arr[i].table = _p(tab, C.c_double)
    return arr


def pattern(K, te, tr, dirichlet_row=None):
    rp = C.POINTER(C.c_int)(); ci = C.POINTER(C.c_int)()
    flags = None if dirichlet_row is None else np.ascontiguousarray(dirichlet_row, dtype=np.uint8)
    nnz = lib().orc_pattern(C.c_size_t(K), C.byref(te.c), C.byref(tr.c),
                            None if flags is None else _p(flags, C.c_ubyte), C.byref(rp), C.byref(ci))
    return _take(rp, te.total + 1, np.int32), _take(ci, nnz, np.int32)


def assemble_bilinear(vertices, cells, te, tr, quad, terms, dirichlet_row, composite, rowptr, colind, val=None):
    vertices = _f64(vertices); cells = _u32(cells)
    keep = []
    arr = _terms(terms, keep)
    flags = None if dirichlet_row is None else np.ascontiguousarray(dirichlet_row, dtype=np.uint8)
    if val is None:
        val = np.zeros(colind.shape[0])
    rowptr = _i32(rowptr); colind = _i32(colind)
    lib().orc_assemble_bilinear(te.dim, _p(vertices, C.c_double), _p(cells, C.c_uint),
                                C.c_size_t(cells.shape[0]), C.byref(te.c), C.byref(tr.c),
                                quad.encode(), arr, len(terms),
                                None if flags is None else _p(flags, C.c_ubyte), int(composite),
                                _p(rowptr, C.c_int), _p(colind, C.c_int), _p(val, C.c_double))
    return val


def assemble_linear(vertices, cells, te, quad, terms, composite, b=None):
    vertices = _f64(vertices); cells = _u32(cells)
    keep = []
    arr = _terms(terms, keep)
    if b is None:
        b = np.zeros(te.total)
    lib().orc_assemble_linear(te.dim, _p(vertices, C.c_double), _p(cells, C.c_uint),
                              C.c_size_t(cells.shape[0]), C.byref(te.c), quad.encode(),
                              arr, len(terms), int(composite), _p(b, C.c_double))
    return b


def quadrature_points(vertices, cells, quad):
    vertices = _f64(vertices); cells = _u32(cells)
    dim = vertices.shape[1]
    nq = quadrature(quad)[1].shape[0]
    xq = np.zeros((cells.shape[0], nq, dim))
    lib().orc_quadrature_points(dim, _p(vertices, C.c_double), _p(cells, C.c_uint),
                                C.c_size_t(cells.shape[0]), quad.encode(), _p(xq, C.c_double))
    return xq


def field_at_quadrature(fe, vertices, cells, dof_map, coef, d, quad):
    vertices = _f64(vertices); cells = _u32(cells); dof_map = _u32(dof_map); coef = _f64(coef)
    dim = vertices.shape[1]
    nq = quadrature(quad)[1].shape[0]
    out = np.zeros((cells.shape[0], nq))
    lib().orc_field_at_quadrature(dim, fe, _p(vertices, C.c_double), _p(cells, C.c_uint),
                                  C.c_size_t(cells.shape[0]), _p(dof_map, C.c_uint),
                                  _p(coef, C.c_double), int(d), quad.encode(), _p(out, C.c_double))
    return out


def integrate(vertices, cells, quad, table):
    vertices = _f64(vertices); cells = _u32(cells); table = _f64(table)
    return lib().orc_integrate(vertices.shape[1], _p(vertices, C.c_double), _p(cells, C.c_uint),
                               C.c_size_t(cells.shape[0]), quad.encode(), _p(table, C.c_double))


def spmv(rowptr, colind, val, x):
    rowptr = _i32(rowptr); colind = _i32(colind); val = _f64(val); x = _f64(x)
    y = np.zeros(rowptr.shape[0] - 1)
    lib().orc_spmv(rowptr.shape[0] - 1, _p(rowptr, C.c_int), _p(colind, C.c_int),
                   _p(val, C.c_double), _p(x, C.c_double), _p(y, C.c_double))
    return y


def cg_jacobi(rowptr, colind, val, b, rtol=1e-8, atol=1e-50, maxits=100000):
    rowptr = _i32(rowptr); colind = _i32(colind); val = _f64(val); b = _f64(b)
    x = np.zeros_like(b); rn = C.c_double(0)
    its = lib().orc_cg_jacobi(b.shape[0], _p(rowptr, C.c_int), _p(colind, C.c_int), _p(val, C.c_double),
                              _p(b, C.c_double), _p(x, C.c_double), C.c_double(rtol), C.c_double(atol),
                              int(maxits), C.byref(rn))
    return x, its, rn.value


def gmres_ilu(rowptr, colind, val, b, restart=1000, levels=2, rtol=1e-8, atol=1e-50, dtol=1e20, maxits=2000):
    rowptr = _i32(rowptr); colind = _i32(colind); val = _f64(val); b = _f64(b)
    x = np.zeros_like(b); rn = C.c_double(0)
    its = lib().orc_gmres_ilu(b.shape[0], _p(rowptr, C.c_int), _p(colind, C.c_int), _p(val, C.c_double),
                              _p(b, C.c_double), _p(x, C.c_double), int(restart), int(levels),
                              C.c_double(rtol), C.c_double(atol), C.c_double(dtol), int(maxits), C.byref(rn))
    return x, its, rn.value


def iluk_pattern(rowptr, colind, levels):
    """(rowptr, colind) of the ILU(levels) factors (sequential row-merge, sum rule)"""
    rowptr = _i32(rowptr); colind = _i32(colind)
    n = rowptr.shape[0] - 1
    out_rp = np.zeros(n + 1, dtype=np.int32)
    out_ci = C.POINTER(C.c_int)()
    lib().orc_iluk_pattern.restype = C.c_size_t
    nnz = lib().orc_iluk_pattern(n, _p(rowptr, C.c_int), _p(colind, C.c_int), int(levels), _p(out_rp, C.c_int), C.byref(out_ci))
    return out_rp, _take(out_ci, nnz, np.int32)


def iluk(rowptr, colind, val, levels):
    """(rowptr, colind, lu) of the ILU(levels) factors: L (unit diagonal not stored) left of the diagonal, U from it on"""
    rowptr = _i32(rowptr); colind = _i32(colind); val = _f64(val)
    n = rowptr.shape[0] - 1
    out_rp = np.zeros(n + 1, dtype=np.int32)
    out_ci = C.POINTER(C.c_int)(); out_lu = C.POINTER(C.c_double)()
    lib().orc_iluk.restype = C.c_size_t
    nnz = lib().orc_iluk(n, _p(rowptr, C.c_int), _p(colind, C.c_int), _p(val, C.c_double), int(levels),
                         _p(out_rp, C.c_int), C.byref(out_ci), C.byref(out_lu))
    return out_rp, _take(out_ci, nnz, np.int32), _take(out_lu, nnz, np.float64)


def ilu0(rowptr, colind, val):
    rowptr = _i32(rowptr); colind = _i32(colind); val = _f64(val)
    lu = np.zeros_like(val)
    st = lib().orc_ilu0(rowptr.shape[0] - 1, _p(rowptr, C.c_int), _p(colind, C.c_int),
                        _p(val, C.c_double), _p(lu, C.c_double))
    if st < 0:
        raise RuntimeError("matrix lacks a structural diagonal")
    return lu
