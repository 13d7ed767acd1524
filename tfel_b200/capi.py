"""ctypes binding of libtfel_cuda.so (include/tfel_cuda.h) and thin RAII-style wrappers.

This is the Python face of the C ABI used by tests/ and bench.py; the C++14 face with the
reference's spellings lives in include/tfel/. There is no CPU path: loading fails loudly when the
library has not been built, and every call fails when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libtfel_cuda.so")

FE = {"p1": 1, "p1b": 2, "p2": 3}
QUAD = {"qf1pT": 1, "qf2pT": 2, "qf5pT": 3, "qf1pTlump": 4,
        "qf1pTet": 10, "qf4pTet": 11, "qf5pTet": 12,
        "qfSym1pTet": 13, "qfSym4pTet": 14, "qfSym10pTet": 15, "qfSym20pTet": 16,
        "qfSym35pTet": 17, "qfSym56pTet": 18}
FACTOR_FIELD, FACTOR_TABLE, FACTOR_PROGRAM = 1, 2, 3
METHOD = {"cg": 1, "gmres": 2}
PRECOND = {"none": 0, "jacobi": 1, "block_jacobi": 2, "ilu0": 3, "ilu": 3}
SCATTER = {"rows": 0, "colored": 1, "patch": 2, "auto": 3, "tile": 4}
SPMV = {"auto": 0, "vector": 1, "stream": 2}


class TfelCudaError(RuntimeError):
    pass


def fe_tabulate(dim, fe, x_hat):
    """reference-element basis values and gradients at one point (host only): [nd][dim + 1]"""
    x = np.ascontiguousarray(x_hat, dtype=np.float64)
    out = np.zeros((10, dim + 1))
    n = C.c_int()
    _ck(lib().tfelcu_fe_tabulate(int(dim), FE[fe] if isinstance(fe, str) else int(fe), x.ctypes.data_as(C.c_void_p),
                                 out.ctypes.data_as(C.c_void_p), C.byref(n)))
    return out[:n.value].copy()


class Op(C.Structure):
    _fields_ = [("op", C.c_int32), ("pad", C.c_int32), ("imm", C.c_double)]


class Factor(C.Structure):
    _fields_ = [("kind", C.c_int32), ("d", C.c_int32), ("fes", C.c_void_p), ("data", C.c_void_p),
                ("ops", C.POINTER(Op)), ("n_ops", C.c_int32), ("pad", C.c_int32), ("fn", C.c_void_p), ("closure", C.c_void_p)]


class Term(C.Structure):
    _fields_ = [("test_comp", C.c_int32), ("test_d", C.c_int32), ("trial_comp", C.c_int32),
                ("trial_d", C.c_int32), ("c", C.c_double), ("n_factors", C.c_int32),
                ("factor", C.c_int32 * 5)]


class SolverParams(C.Structure):
    _fields_ = [("method", C.c_int32), ("precond", C.c_int32), ("maxits", C.c_uint32),
                ("restart", C.c_uint32), ("rtol", C.c_double), ("atol", C.c_double),
                ("dtol", C.c_double), ("block_size", C.c_int32), ("check_every", C.c_int32),
                ("spmv_kernel", C.c_int32), ("profile_every", C.c_int32),
                ("ilu_levels", C.c_int32), ("pad", C.c_int32)]


class SolverReport(C.Structure):
    _fields_ = [("its", C.c_uint32), ("converged", C.c_int32), ("rnorm", C.c_double),
                ("bnorm", C.c_double), ("t_setup_s", C.c_double), ("t_solve_s", C.c_double),
                ("n_launches", C.c_uint64), ("n_spmv", C.c_uint64), ("spmv_avg_ms", C.c_double),
                ("spmv_samples", C.c_uint32), ("restart_used", C.c_uint32), ("ilu_nnz", C.c_uint64),
                ("ilu_levels", C.c_int32), ("ilu_zero_pivot_row", C.c_int32), ("t_precond_setup_s", C.c_double),
                ("t_symbolic_s", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/tfel_cuda.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "tfelcu_last_error", "tfelcu_version", "tfelcu_ctx_create", "tfelcu_nccl_unique_id",
    "tfelcu_ctx_create_distributed", "tfelcu_ctx_destroy", "tfelcu_ctx_set_stream", "tfelcu_ctx_stream",
    "tfelcu_ctx_synchronize", "tfelcu_ctx_rank", "tfelcu_ctx_launch_count", "tfelcu_ctx_p2p_enabled", "tfelcu_ctx_p2p_bench", "tfelcu_ctx_fp64_peak",
    "tfelcu_mesh_create", "tfelcu_mesh_gen_square", "tfelcu_mesh_gen_cube", "tfelcu_mesh_destroy",
    "tfelcu_mesh_info", "tfelcu_mesh_download", "tfelcu_mesh_geometry", "tfelcu_mesh_boundary",
    "tfelcu_mesh_boundary_download", "tfelcu_mesh_neighbours",
    "tfelcu_fes_create", "tfelcu_fes_destroy", "tfelcu_fes_info", "tfelcu_fes_kind_offsets", "tfelcu_fes_download_dof_map",
    "tfelcu_fes_download_g2l", "tfelcu_fes_dof_coordinates", "tfelcu_fes_boundary_dofs",
    "tfelcu_fes_boundary_dofs_get", "tfelcu_fes_add_dirichlet", "tfelcu_fes_add_dirichlet_const",
    "tfelcu_fes_dirichlet_count", "tfelcu_fes_dirichlet_get", "tfelcu_fes_vertex_values", "tfelcu_fes_cell_values", "tfelcu_fe_tabulate",
    "tfelcu_matrix_create", "tfelcu_matrix_create_rows", "tfelcu_matrix_local_rows", "tfelcu_matrix_destroy", "tfelcu_matrix_clear", "tfelcu_matrix_info",
    "tfelcu_matrix_set_scatter", "tfelcu_matrix_assemble", "tfelcu_matrix_download_csr",
    "tfelcu_matrix_color_count",
    "tfelcu_vector_create", "tfelcu_vector_create_rows", "tfelcu_vector_destroy", "tfelcu_vector_clear", "tfelcu_vector_size",
    "tfelcu_vector_assemble", "tfelcu_vector_set_scatter", "tfelcu_vector_download", "tfelcu_vector_upload", "tfelcu_vector_device_ptr",
    "tfelcu_integrate", "tfelcu_quadrature_size", "tfelcu_mesh_quadrature_points",
    "tfelcu_solver_create", "tfelcu_solver_destroy", "tfelcu_solver_set_operator",
    "tfelcu_solver_set_operator_csr", "tfelcu_solver_solve", "tfelcu_matrix_solve", "tfelcu_matrix_make_rhs",
    "tfelcu_solver_gather", "tfelcu_solver_spmv", "tfelcu_solver_spmv_bench",
    "tfelcu_solver_ilu_info", "tfelcu_solver_ilu_download",
    "tfelcu_array_create", "tfelcu_array_destroy", "tfelcu_array_copy", "tfelcu_array_lincomb",
]

_lib = None


def lib():
    """The loaded library. Raises when it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TfelCudaError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(make -C tfel_b200/csrc); tfel_b200 has no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.tfelcu_last_error.restype = C.c_char_p
        L.tfelcu_version.restype = C.c_char_p
        L.tfelcu_ctx_stream.restype = C.c_void_p
        L.tfelcu_ctx_stream.argtypes = [C.c_void_p]
        L.tfelcu_ctx_launch_count.restype = C.c_uint64
        L.tfelcu_ctx_launch_count.argtypes = [C.c_void_p]
        L.tfelcu_vector_device_ptr.restype = C.c_void_p
        L.tfelcu_vector_device_ptr.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ck(code):
    if code != 0:
        raise TfelCudaError(lib().tfelcu_last_error().decode())


def _ptr(a):
    """void* of a numpy array, a torch tensor (host or device), an int address, or None."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def _sz(n):
    return C.c_size_t(int(n))


class Context:
    def __init__(self, device=0, rank=None, n_ranks=None, nccl_id=None):
        self.h = C.c_void_p()
        if n_ranks is None or n_ranks == 1 and nccl_id is None:
            _ck(lib().tfelcu_ctx_create(int(device), C.byref(self.h)))
        else:
            buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
            _ck(lib().tfelcu_ctx_create_distributed(int(device), int(rank), int(n_ranks), buf, C.byref(self.h)))

    @staticmethod
    def nccl_unique_id():
        buf = (C.c_char * 128)()
        _ck(lib().tfelcu_nccl_unique_id(buf))
        return bytes(buf)

    def close(self):
        if self.h:
            lib().tfelcu_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def synchronize(self):
        _ck(lib().tfelcu_ctx_synchronize(self.h))

    def set_stream(self, cuda_stream):
        _ck(lib().tfelcu_ctx_set_stream(self.h, C.c_void_p(int(cuda_stream))))

    @property
    def stream(self):
        return lib().tfelcu_ctx_stream(self.h)

    @property
    def launches(self):
        return int(lib().tfelcu_ctx_launch_count(self.h))

    @property
    def p2p(self):
        return bool(lib().tfelcu_ctx_p2p_enabled(self.h))

    def fp64_peak(self):
        """measured FP64 FMA throughput of the device, TFLOP/s"""
        t = C.c_double()
        _ck(lib().tfelcu_ctx_fp64_peak(self.h, C.byref(t)))
        return t.value

    def p2p_bench(self, reps=1000):
        us = C.c_double()
        _ck(lib().tfelcu_ctx_p2p_bench(self.h, int(reps), C.byref(us)))
        return us.value


class Mesh:
    """mesh<cell> / fe_mesh<cell> (src/core/mesh.hpp:142-158, 355-364) on device."""

    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h
        dim = C.c_int(); V = C.c_size_t(); K = C.c_size_t()
        _ck(lib().tfelcu_mesh_info(h, C.byref(dim), C.byref(V), C.byref(K)))
        self.dim, self.n_vertices, self.n_cells = dim.value, V.value, K.value

    @classmethod
    def create(cls, ctx, vertices, cells):
        vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        cells = np.ascontiguousarray(cells, dtype=np.uint32)
        h = C.c_void_p()
        _ck(lib().tfelcu_mesh_create(ctx.h, int(vertices.shape[1]), _ptr(vertices), _sz(vertices.shape[0]),
                                     _ptr(cells), _sz(cells.shape[0]), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def square(cls, ctx, x1, x2, n1, n2):
        h = C.c_void_p()
        _ck(lib().tfelcu_mesh_gen_square(ctx.h, C.c_double(x1), C.c_double(x2), C.c_uint32(n1), C.c_uint32(n2), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def cube(cls, ctx, x1, x2, x3, n1, n2, n3):
        h = C.c_void_p()
        _ck(lib().tfelcu_mesh_gen_cube(ctx.h, C.c_double(x1), C.c_double(x2), C.c_double(x3),
                                       C.c_uint32(n1), C.c_uint32(n2), C.c_uint32(n3), C.byref(h)))
        return cls(ctx, h)

    def close(self):
        if self.h:
            lib().tfelcu_mesh_destroy(self.h)
            self.h = C.c_void_p()

    def download(self):
        v = np.zeros((self.n_vertices, self.dim)); c = np.zeros((self.n_cells, self.dim + 1), dtype=np.uint32)
        _ck(lib().tfelcu_mesh_download(self.h, _ptr(v), _ptr(c)))
        return v, c

    def geometry(self):
        vol = np.zeros(self.n_cells); jmt = np.zeros((self.n_cells, self.dim, self.dim))
        _ck(lib().tfelcu_mesh_geometry(self.h, _ptr(vol), _ptr(jmt)))
        return vol, jmt

    def quadrature_points(self, quad):
        nq = C.c_int(); dim = C.c_int()
        _ck(lib().tfelcu_quadrature_size(QUAD[quad], C.byref(nq), C.byref(dim)))
        xq = np.zeros((self.n_cells, nq.value, self.dim))
        _ck(lib().tfelcu_mesh_quadrature_points(self.h, QUAD[quad], _ptr(xq)))
        return xq

    def boundary(self):
        F = C.c_size_t()
        _ck(lib().tfelcu_mesh_boundary(self.h, C.byref(F)))
        el = np.zeros((F.value, self.dim), dtype=np.uint32)
        pc = np.zeros(F.value, dtype=np.uint32); ps = np.zeros(F.value, dtype=np.uint32)
        _ck(lib().tfelcu_mesh_boundary_download(self.h, _ptr(el), _ptr(pc), _ptr(ps)))
        return el, pc, ps


    def neighbours(self):
        """mesh<cell>::compute_cell_neighbours (mesh.hpp:301-332): [K][dim + 1] int32, -1 on the boundary"""
        nb = np.zeros((self.n_cells, self.dim + 1), dtype=np.int32)
        _ck(lib().tfelcu_mesh_neighbours(self.h, _ptr(nb)))
        return nb


class Space:
    """finite_element_space<fe> (src/core/fes.hpp:18-82)."""

    def __init__(self, mesh, fe):
        self.mesh, self.ctx = mesh, mesh.ctx
        self.h = C.c_void_p()
        _ck(lib().tfelcu_fes_create(self.ctx.h, mesh.h, FE[fe] if isinstance(fe, str) else int(fe), C.byref(self.h)))
        fe_id = C.c_int(); n = C.c_size_t(); nd = C.c_int()
        _ck(lib().tfelcu_fes_info(self.h, C.byref(fe_id), C.byref(n), C.byref(nd)))
        self.fe, self.n_dof, self.nd = fe_id.value, n.value, nd.value

    def close(self):
        if self.h:
            lib().tfelcu_fes_destroy(self.h)
            self.h = C.c_void_p()

    def kind_offsets(self):
        """first dof of the vertex / edge / face / cell dofs, then n_dof"""
        out = (C.c_size_t * 5)()
        _ck(lib().tfelcu_fes_kind_offsets(self.h, out))
        return [int(v) for v in out]

    def dof_map(self):
        out = np.zeros((self.mesh.n_cells, self.nd), dtype=np.uint32)
        _ck(lib().tfelcu_fes_download_dof_map(self.h, _ptr(out)))
        return out

    def g2l(self):
        out = np.zeros((self.n_dof, 2), dtype=np.uint32)
        _ck(lib().tfelcu_fes_download_g2l(self.h, _ptr(out)))
        return out

    def dof_coordinates(self, dofs=None):
        dofs = np.arange(self.n_dof, dtype=np.uint32) if dofs is None else np.ascontiguousarray(dofs, dtype=np.uint32)
        x = np.zeros((dofs.shape[0], self.mesh.dim))
        _ck(lib().tfelcu_fes_dof_coordinates(self.h, _ptr(dofs), _sz(dofs.shape[0]), _ptr(x)))
        return x

    def vertex_values(self, coefficients):
        """to_mesh_vertex_data (fes.hpp:483-504)"""
        c = np.ascontiguousarray(coefficients, dtype=np.float64)
        out = np.zeros(self.mesh.n_vertices)
        _ck(lib().tfelcu_fes_vertex_values(self.h, _ptr(c), _ptr(out)))
        return out

    def cell_values(self, coefficients):
        """to_mesh_cell_data (fes.hpp:464-481): the discrete function at the barycentre of every cell"""
        c = np.ascontiguousarray(coefficients, dtype=np.float64)
        out = np.zeros(self.mesh.n_cells)
        _ck(lib().tfelcu_fes_cell_values(self.h, _ptr(c), _ptr(out)))
        return out

    def boundary_dofs(self, bcells=None, with_coordinates=False):
        """dofs add_dirichlet_boundary would touch (fes.hpp:100-122); bcells None = whole boundary."""
        n = C.c_size_t()
        if bcells is None:
            _ck(lib().tfelcu_fes_boundary_dofs(self.h, None, _sz(0), 0, C.byref(n)))
        else:
            bcells = np.ascontiguousarray(bcells, dtype=np.uint32)
            _ck(lib().tfelcu_fes_boundary_dofs(self.h, _ptr(bcells), _sz(bcells.shape[0]), int(bcells.shape[1]), C.byref(n)))
        dofs = np.zeros(n.value, dtype=np.uint32)
        x = np.zeros((n.value, self.mesh.dim)) if with_coordinates else None
        _ck(lib().tfelcu_fes_boundary_dofs_get(self.h, _ptr(dofs), _ptr(x)))
        return (dofs, x) if with_coordinates else dofs

    def add_dirichlet(self, dofs, values):
        dofs = np.ascontiguousarray(dofs, dtype=np.uint32)
        values = np.ascontiguousarray(np.broadcast_to(values, dofs.shape), dtype=np.float64)
        _ck(lib().tfelcu_fes_add_dirichlet(self.h, _ptr(dofs), _ptr(values), _sz(dofs.shape[0])))

    def add_dirichlet_boundary(self, value=0.0, bcells=None):
        """add_dirichlet_boundary(dm, value | function) (fes.hpp:100-149)."""
        if callable(value):
            dofs, x = self.boundary_dofs(bcells, with_coordinates=True)
            self.add_dirichlet(dofs, value(x))
        else:
            n = C.c_size_t()
            if bcells is None:
                _ck(lib().tfelcu_fes_boundary_dofs(self.h, None, _sz(0), 0, C.byref(n)))
            else:
                bcells = np.ascontiguousarray(bcells, dtype=np.uint32)
                _ck(lib().tfelcu_fes_boundary_dofs(self.h, _ptr(bcells), _sz(bcells.shape[0]), int(bcells.shape[1]), C.byref(n)))
            _ck(lib().tfelcu_fes_add_dirichlet_const(self.h, C.c_double(float(value))))

    def dirichlet(self):
        n = C.c_size_t()
        _ck(lib().tfelcu_fes_dirichlet_count(self.h, C.byref(n)))
        dofs = np.zeros(n.value, dtype=np.uint32); vals = np.zeros(n.value)
        if n.value:
            _ck(lib().tfelcu_fes_dirichlet_get(self.h, _ptr(dofs), _ptr(vals)))
        return dofs, vals


class _Integrand:
    """Sum-of-products integrand lowered to the C structs (keeps the referenced buffers alive)."""

    def __init__(self, terms, fields=None, spaces=None):
        """terms: dicts {test:(comp,d), trial:(comp,d)?, c:float, factors:[("x", X|array|float) | ("field", name, d)]};
        fields: name -> (Space, coefficient array (numpy) or device pointer)."""
        self.keep = []
        factors = []
        index = {}

        def factor_index(f):
            if f[0] == "x":
                key = ("x", id(f[1]))
                if key in index:
                    return index[key]
                fa = Factor()
                v = f[1]
                if hasattr(v, "program"):
                    prog = v.program()
                    ops = (Op * len(prog))()
                    for i, (op, imm) in enumerate(prog):
                        ops[i].op = int(op); ops[i].imm = float(imm)
                    self.keep.append(ops)
                    fa.kind = FACTOR_PROGRAM; fa.ops = C.cast(ops, C.POINTER(Op)); fa.n_ops = len(prog)
                elif isinstance(v, (int, float)):
                    ops = (Op * 1)()
                    ops[0].op = 0; ops[0].imm = float(v)
                    self.keep.append(ops)
                    fa.kind = FACTOR_PROGRAM; fa.ops = C.cast(ops, C.POINTER(Op)); fa.n_ops = 1
                else:   # table [K][nq]: numpy array or device tensor
                    if isinstance(v, np.ndarray):
                        v = np.ascontiguousarray(v, dtype=np.float64)
                    self.keep.append(v)
                    fa.kind = FACTOR_TABLE; fa.data = _ptr(v)
            else:
                key = ("field", f[1], f[2])
                if key in index:
                    return index[key]
                space, coef = fields[f[1]]
                if isinstance(coef, np.ndarray):
                    coef = np.ascontiguousarray(coef, dtype=np.float64)
                self.keep.append(coef)
                fa = Factor()
                fa.kind = FACTOR_FIELD; fa.d = int(f[2]); fa.fes = space.h; fa.data = _ptr(coef)
            factors.append(fa)
            index[key] = len(factors) - 1
            return index[key]

        cterms = (Term * max(len(terms), 1))()
        for i, t in enumerate(terms):
            ct = cterms[i]
            ct.test_comp, ct.test_d = t["test"]
            ct.trial_comp, ct.trial_d = t.get("trial", (-1, 0))
            ct.c = float(t.get("c", 1.0))
            fl = t.get("factors", [])
            ct.n_factors = len(fl)
            for j, f in enumerate(fl):
                ct.factor[j] = factor_index(f)
        self.terms, self.n_terms = cterms, len(terms)
        self.factors = (Factor * max(len(factors), 1))(*factors)
        self.n_factors = len(factors)


def integrand(terms, fields=None):
    """An integrand lowered once to the C structs of the ABI, for `+=` calls that repeat (time / Newton loops, bench.py):
    Matrix.assemble / Vector.assemble accept it in place of the list of terms."""
    return _Integrand(terms, fields)


class Matrix:
    """bilinear_form<te, tr> and its sparse_matrix (bilinear_form.hpp:9-19, solver.hpp:236-283)."""

    def __init__(self, test, trial=None, scatter=None, rows=None):
        """rows: None (all rows) or a list of (begin, end) global row segments this rank owns"""
        test = list(test) if isinstance(test, (list, tuple)) else [test]
        trial = test if trial is None else (list(trial) if isinstance(trial, (list, tuple)) else [trial])
        self.test, self.trial, self.ctx = test, trial, test[0].ctx
        self.h = C.c_void_p()
        te = (C.c_void_p * len(test))(*[s.h for s in test])
        tr = (C.c_void_p * len(trial))(*[s.h for s in trial])
        if rows is None:
            _ck(lib().tfelcu_matrix_create(self.ctx.h, len(test), te, len(trial), tr, C.byref(self.h)))
        else:
            rb = (C.c_size_t * len(rows))(*[int(r[0]) for r in rows]); re = (C.c_size_t * len(rows))(*[int(r[1]) for r in rows])
            _ck(lib().tfelcu_matrix_create_rows(self.ctx.h, len(test), te, len(trial), tr, len(rows), rb, re, C.byref(self.h)))
        if scatter is not None:
            _ck(lib().tfelcu_matrix_set_scatter(self.h, SCATTER[scatter]))
        self._info()

    def _info(self):
        nr = C.c_size_t(); nc = C.c_size_t(); nnz = C.c_size_t()
        _ck(lib().tfelcu_matrix_info(self.h, C.byref(nr), C.byref(nc), C.byref(nnz)))
        self.n_rows, self.n_cols, self.nnz = nr.value, nc.value, nnz.value
        nl = C.c_size_t()
        _ck(lib().tfelcu_matrix_local_rows(self.h, C.byref(nl)))
        self.n_local = nl.value

    def close(self):
        if self.h:
            lib().tfelcu_matrix_destroy(self.h)
            self.h = C.c_void_p()

    def clear(self, refresh_info=True):
        _ck(lib().tfelcu_matrix_clear(self.h))
        if refresh_info:      # the pattern (nnz) changes with the Dirichlet sets of the test spaces
            self._info()

    def assemble(self, quad, terms, fields=None):
        """a += integrate<quad>(expr, mesh); terms: list of term dicts or a prepared capi.integrand(...)"""
        ig = terms if isinstance(terms, _Integrand) else _Integrand(terms, fields)
        _ck(lib().tfelcu_matrix_assemble(self.h, QUAD[quad], ig.factors, ig.n_factors, ig.terms, ig.n_terms))

    def csr(self, values=True):
        self._info()
        rowptr = np.zeros(self.n_local + 1, dtype=np.int32); colind = np.zeros(self.nnz, dtype=np.int32)
        val = np.zeros(self.nnz) if values else None
        _ck(lib().tfelcu_matrix_download_csr(self.h, _ptr(rowptr), _ptr(colind), _ptr(val)))
        return (rowptr, colind, val) if values else (rowptr, colind)

    def color_count(self):
        n = C.c_int()
        _ck(lib().tfelcu_matrix_color_count(self.h, C.byref(n)))
        return n.value

    def make_rhs(self, b):
        out = np.zeros(self.n_local)
        _ck(lib().tfelcu_matrix_make_rhs(self.h, b.h, _ptr(out)))
        return out

    def solve(self, b, solver, out=None):
        """a.solve(b, s) (bilinear_form.hpp:122-157) -> (x, report)"""
        x = np.zeros(self.n_cols) if out is None else out
        rep = SolverReport()
        _ck(lib().tfelcu_matrix_solve(self.h, b.h, solver.h, _ptr(x), C.byref(rep)))
        return x, rep.as_dict()


class Vector:
    """linear_form<fes> (linear_form.hpp:9-20)."""

    def __init__(self, test, rows=None, scatter=None):
        test = list(test) if isinstance(test, (list, tuple)) else [test]
        self.test, self.ctx = test, test[0].ctx
        self.h = C.c_void_p()
        te = (C.c_void_p * len(test))(*[s.h for s in test])
        if rows is None:
            _ck(lib().tfelcu_vector_create(self.ctx.h, len(test), te, C.byref(self.h)))
        else:
            rb = (C.c_size_t * len(rows))(*[int(r[0]) for r in rows]); re = (C.c_size_t * len(rows))(*[int(r[1]) for r in rows])
            _ck(lib().tfelcu_vector_create_rows(self.ctx.h, len(test), te, len(rows), rb, re, C.byref(self.h)))
        if scatter is not None and scatter != "colored":   # a linear form has no colour-by-colour variant
            _ck(lib().tfelcu_vector_set_scatter(self.h, SCATTER[scatter]))
        n = C.c_size_t()
        _ck(lib().tfelcu_vector_size(self.h, C.byref(n)))
        self.n = n.value

    def close(self):
        if self.h:
            lib().tfelcu_vector_destroy(self.h)
            self.h = C.c_void_p()

    def clear(self):
        _ck(lib().tfelcu_vector_clear(self.h))

    def assemble(self, quad, terms, fields=None):
        ig = terms if isinstance(terms, _Integrand) else _Integrand(terms, fields)
        _ck(lib().tfelcu_vector_assemble(self.h, QUAD[quad], ig.factors, ig.n_factors, ig.terms, ig.n_terms))

    def download(self):
        out = np.zeros(self.n)
        _ck(lib().tfelcu_vector_download(self.h, _ptr(out)))
        return out

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape[0] == self.n
        _ck(lib().tfelcu_vector_upload(self.h, _ptr(a)))

    @property
    def device_ptr(self):
        return lib().tfelcu_vector_device_ptr(self.h)


def integrate(mesh, quad, terms, fields=None):
    """rank-0 integrate<quad>(expr, mesh) (form.hpp:87-139); terms carry only c and factors."""
    ig = _Integrand([dict(test=(0, 0), trial=(0, 0), **{k: v for k, v in t.items() if k in ("c", "factors")}) for t in terms], fields)
    out = C.c_double()
    _ck(lib().tfelcu_integrate(mesh.ctx.h, mesh.h, QUAD[quad], ig.factors, ig.n_factors, ig.terms, ig.n_terms, C.byref(out)))
    return out.value


class Solver:
    """solver::cuda::* behind solver::basic_solver (solver.hpp:29-40); same dictionary keys as
    solver::petsc::gmres_ilu (solver.hpp:125-133) plus method / precond."""

    def __init__(self, ctx, method="cg", precond="jacobi", maxits=2000, restart=1000, rtol=1e-8, atol=1e-50,
                 dtol=1e20, block_size=0, check_every=0, spmv="auto", profile_every=0, ilufill=0):
        self.ctx = ctx
        if precond == "ilu0":
            ilufill = 0
        p = SolverParams(METHOD[method], PRECOND[precond], int(maxits), int(restart), float(rtol), float(atol),
                         float(dtol), int(block_size), int(check_every), SPMV[spmv], int(profile_every), int(ilufill), 0)
        self.h = C.c_void_p()
        _ck(lib().tfelcu_solver_create(ctx.h, C.byref(p), C.byref(self.h)))
        self.n = 0

    def close(self):
        if self.h:
            lib().tfelcu_solver_destroy(self.h)
            self.h = C.c_void_p()

    def set_operator(self, a):
        """a: Matrix (device resident, zero copy) or (rowptr, colind, val) host CSR."""
        if isinstance(a, Matrix):
            _ck(lib().tfelcu_solver_set_operator(self.h, a.h))
            self.n = a.n_local
            self.n_global = a.n_rows
        else:
            rowptr, colind, val = a
            rowptr = np.ascontiguousarray(rowptr, dtype=np.int32); colind = np.ascontiguousarray(colind, dtype=np.int32)
            val = np.ascontiguousarray(val, dtype=np.float64)
            self.n = rowptr.shape[0] - 1
            _ck(lib().tfelcu_solver_set_operator_csr(self.h, _sz(self.n), _ptr(rowptr), _ptr(colind), _ptr(val)))

    def solve(self, rhs, x=None):
        if isinstance(rhs, np.ndarray):
            rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        if x is None:
            x = np.zeros(self.n)
        rep = SolverReport()
        _ck(lib().tfelcu_solver_solve(self.h, _ptr(rhs), _ptr(x), C.byref(rep)))
        return x, rep.as_dict()

    def spmv(self, x, y=None):
        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(x, dtype=np.float64)
        if y is None:
            y = np.zeros(self.n)
        _ck(lib().tfelcu_solver_spmv(self.h, _ptr(x), _ptr(y)))
        return y

    def gather(self, x_local, out=None):
        """owned entries of every rank -> full vector in reference numbering (on every rank)"""
        if isinstance(x_local, np.ndarray):
            x_local = np.ascontiguousarray(x_local, dtype=np.float64)
        if out is None:
            out = np.zeros(self.n_global)
        _ck(lib().tfelcu_solver_gather(self.h, _ptr(x_local), _ptr(out)))
        return out

    def ilu_factors(self):
        """(rowptr int64 [n + 1], colind int32, lu) of the ILU(ilufill) factors the solver holds (parity checks)"""
        n = C.c_size_t(); nnz = C.c_size_t(); lv = C.c_int()
        _ck(lib().tfelcu_solver_ilu_info(self.h, C.byref(n), C.byref(nnz), C.byref(lv)))
        rp = np.zeros(n.value + 1, dtype=np.int64); ci = np.zeros(nnz.value, dtype=np.int32); lu = np.zeros(nnz.value)
        _ck(lib().tfelcu_solver_ilu_download(self.h, _ptr(rp), _ptr(ci), _ptr(lu)))
        return rp, ci, lu

    def spmv_bench(self, reps=20):
        ms = C.c_double(); nbytes = C.c_double()
        _ck(lib().tfelcu_solver_spmv_bench(self.h, int(reps), C.byref(ms), C.byref(nbytes)))
        return ms.value, nbytes.value
