"""Problem definitions shared by the oracle tests (CPU) and the CUDA parity tests (GPU).

Each problem mirrors one set-up of oracle/ref_driver.cpp, which in turn follows the
reference's own programs (README.md:66-120, test/tetrahedron_cube.cpp,
test/stokes_2d_p2_p1.cpp, test/navier_stokes_2d_p2_p1.cpp, test/composite_fe.cpp).
Integrands are written in the sum-of-products normal form
    sum_t  c_t * prod(factors_t) * D^{a_t} psi_i * D^{b_t} phi_j
where a factor is a function of x (an `X` expression: evaluable with numpy for the oracle,
lowerable to the postfix program the C-ABI takes) or a finite element field.
"""


import numpy as np


class X:
    """Tiny expression over space coordinates: numpy evaluation + postfix program."""
    CONST, COORD, ADD, SUB, MUL, DIV, NEG, SQRT, EXP, SIN, COS, POW, LT, ABS = range(14)

    def __init__(self, op, args=(), imm=0.0):
        self.op, self.args, self.imm = op, args, imm

    @staticmethod
    def c(v):
        return v if isinstance(v, X) else X(X.CONST, (), float(v))

    @staticmethod
    def x(i):
        return X(X.COORD, (), float(i))

    def __add__(self, o): return X(X.ADD, (self, X.c(o)))
    def __radd__(self, o): return X(X.ADD, (X.c(o), self))
    def __sub__(self, o): return X(X.SUB, (self, X.c(o)))
    def __rsub__(self, o): return X(X.SUB, (X.c(o), self))
    def __mul__(self, o): return X(X.MUL, (self, X.c(o)))
    def __rmul__(self, o): return X(X.MUL, (X.c(o), self))
    def __truediv__(self, o): return X(X.DIV, (self, X.c(o)))
    def __rtruediv__(self, o): return X(X.DIV, (X.c(o), self))
    def __neg__(self): return X(X.NEG, (self,))
    def __pow__(self, o): return X(X.POW, (self, X.c(o)))
    def lt(self, o): return X(X.LT, (self, X.c(o)))

    @staticmethod
    def sqrt(a): return X(X.SQRT, (X.c(a),))
    @staticmethod
    def exp(a): return X(X.EXP, (X.c(a),))
    @staticmethod
    def sin(a): return X(X.SIN, (X.c(a),))
    @staticmethod
    def cos(a): return X(X.COS, (X.c(a),))

    def eval(self, x):
        """x: array [..., dim] -> array [...]"""
        x = np.asarray(x, dtype=np.float64)
        a = [t.eval(x) for t in self.args]
        o = self.op
        if o == X.CONST: return np.full(x.shape[:-1], self.imm)
        if o == X.COORD: return x[..., int(self.imm)].copy()
        if o == X.ADD: return a[0] + a[1]
        if o == X.SUB: return a[0] - a[1]
        if o == X.MUL: return a[0] * a[1]
        if o == X.DIV: return a[0] / a[1]
        if o == X.NEG: return -a[0]
        if o == X.SQRT: return np.sqrt(a[0])
        if o == X.EXP: return np.exp(a[0])
        if o == X.SIN: return np.sin(a[0])
        if o == X.COS: return np.cos(a[0])
        if o == X.POW: return np.power(a[0], a[1])
        if o == X.LT: return (a[0] < a[1]).astype(np.float64)
        if o == X.ABS: return np.abs(a[0])
        raise ValueError(o)

    def program(self):
        """Postfix list of (opcode, immediate)."""
        out = []
        for t in self.args:
            out += t.program()
        out.append((self.op, self.imm))
        return out


x0, x1, x2 = X.x(0), X.x(1), X.x(2)


def lap_terms(dim, comp_te=0, comp_tr=0, c=1.0):
    return [dict(test=(comp_te, d), trial=(comp_tr, d), c=c) for d in range(1, dim + 1)]


class Problem:
    def __init__(self, name, golden, mesh, comps, quad, composite, bilinear, linear,
                 dirichlet, fields=None):
        self.name, self.golden, self.mesh = name, golden, mesh
        self.comps, self.quad, self.composite = comps, quad, composite
        self.bilinear, self.linear, self.dirichlet = bilinear, linear, dirichlet
        self.fields = fields or {}
        self.dim = 2 if mesh[0] == "square" else 3

    def with_n(self, n):
        import copy
        p = copy.copy(self)
        p.mesh = (self.mesh[0], n)
        p.golden = None
        return p


# ---- Dirichlet data (host functions of the dof coordinate, like the reference's) ----
def g_zero(x): return np.zeros(x.shape[0])
def g_bc(x): return 1.0 + x[:, 0] + 2.0 * x[:, 1] + x[:, 0] * x[:, 1]
def tet_u_bc(x): return x[:, 0] * x[:, 0] + x[:, 1] * x[:, 1] + x[:, 2] * x[:, 2]
def st_u0_bv(x): return np.where(x[:, 1] < 0.0001, x[:, 0] * (1.0 - x[:, 0]), 0.0)
def st_u1_bv(x): return np.where(x[:, 0] < 0.0001, x[:, 1] * (1.0 - x[:, 1]), 0.0)
def ns_u0_bv(x): return np.where(x[:, 1] < 0.00001, 4.0 * x[:, 0] * (1.0 - x[:, 0]), 0.0)
def ns_u1_bv(x): return np.where(x[:, 0] < 0.00001, 4.0 * x[:, 1] * (1.0 - x[:, 1]), 0.0)


gaussian_src = X.exp(-(X.sqrt((x0 - 0.5) ** 2.0 + (x1 - 0.5) ** 2.0) * X.sqrt((x0 - 0.5) ** 2.0 + (x1 - 0.5) ** 2.0))
                     / (2.0 * 0.1 * 0.1))
c_reac = 1.0 + x0 * x0 + 0.5 * x1


def poisson(fe, n, golden=None):
    return Problem("poisson_%s" % fe, golden, ("square", n), [fe], "qf5pT", False,
                   lap_terms(2), [dict(test=(0, 0), c=1.0)],
                   [dict(comp=0, where="boundary", value=0.0)])


def poissong(fe, n, golden=None):
    return Problem("poissong_%s" % fe, golden, ("square", n), [fe], "qf5pT", False,
                   lap_terms(2) + [dict(test=(0, 0), trial=(0, 0), c=1.0, factors=[("x", c_reac)])],
                   [dict(test=(0, 0), c=1.0, factors=[("x", gaussian_src)])],
                   [dict(comp=0, where="boundary", value=g_bc)])


def tet(fe, n, golden=None):
    quad = "qfSym4pTet" if fe == "p1" else ("qfSym35pTet" if fe == "p1b" else "qfSym10pTet")
    return Problem("tet_%s" % fe, golden, ("cube", n), [fe], quad, False,
                   lap_terms(3), [dict(test=(0, 0), c=-6.0)],
                   [dict(comp=0, where="boundary", value=tet_u_bc)])


def stokes(n, golden=None):
    bil = (lap_terms(2, 0, 0) + lap_terms(2, 1, 1)
           + [dict(test=(0, 1), trial=(2, 0)), dict(test=(1, 2), trial=(2, 0))]    # p (d1 v0 + d2 v1)
           + [dict(test=(2, 0), trial=(0, 1)), dict(test=(2, 0), trial=(1, 2))])   # q (d1 u0 + d2 u1)
    lin = [dict(test=(0, 0), c=0.0), dict(test=(1, 0), c=0.0)]
    return Problem("stokes", golden, ("square", n), ["p2", "p2", "p1"], "qf5pT", True, bil, lin,
                   [dict(comp=0, where="boundary", value=st_u0_bv),
                    dict(comp=1, where="boundary", value=st_u1_bv),
                    dict(comp=2, where="pin", value=0.0)])


def ns_prev0(x): return np.sin(3.0 * x[:, 0]) * np.cos(2.0 * x[:, 1]) + 0.25
def ns_prev1(x): return x[:, 0] * x[:, 1] - 0.5 * x[:, 1] * x[:, 1] + 0.1
def ns_old0(x): return 0.5 * x[:, 0] - x[:, 1] * x[:, 1]
def ns_old1(x): return np.cos(x[:, 0] + 2.0 * x[:, 1])


def navier_stokes(n, golden=None):
    dt, re = 0.5, 5000.0
    v0, v1 = ("field", "prev0"), ("field", "prev1")

    def F(f, d=0):
        return (f[0], f[1], d)
    bil = [dict(test=(0, 0), trial=(0, 0)), dict(test=(1, 0), trial=(1, 0))]
    # dt (vn0 d1(v0) w0 + vn0 d1(v1) w1 + vn1 d2(v0) w0 + vn1 d2(v1) w1)
    bil += [dict(test=(0, 0), trial=(0, 0), c=dt, factors=[F(v0, 1)]),
            dict(test=(1, 0), trial=(0, 0), c=dt, factors=[F(v1, 1)]),
            dict(test=(0, 0), trial=(1, 0), c=dt, factors=[F(v0, 2)]),
            dict(test=(1, 0), trial=(1, 0), c=dt, factors=[F(v1, 2)])]
    # dt (v0 d1(vn0) w0 + v0 d1(vn1) w1 + v1 d2(vn0) w0 + v1 d2(vn1) w1)
    bil += [dict(test=(0, 0), trial=(0, 1), c=dt, factors=[F(v0)]),
            dict(test=(1, 0), trial=(1, 1), c=dt, factors=[F(v0)]),
            dict(test=(0, 0), trial=(0, 2), c=dt, factors=[F(v1)]),
            dict(test=(1, 0), trial=(1, 2), c=dt, factors=[F(v1)])]
    bil += lap_terms(2, 0, 0, dt / re) + lap_terms(2, 1, 1, dt / re)
    bil += [dict(test=(0, 1), trial=(2, 0), c=dt), dict(test=(1, 2), trial=(2, 0), c=dt)]
    bil += [dict(test=(2, 0), trial=(0, 1), c=dt), dict(test=(2, 0), trial=(1, 2), c=dt)]
    u0, u1 = ("field", "old0"), ("field", "old1")
    lin = [dict(test=(0, 0), factors=[F(u0)]), dict(test=(1, 0), factors=[F(u1)]),
           dict(test=(0, 0), c=0.0), dict(test=(1, 0), c=0.0),
           dict(test=(0, 0), c=dt, factors=[F(v0), F(v0, 1)]),
           dict(test=(1, 0), c=dt, factors=[F(v0), F(v1, 1)]),
           dict(test=(0, 0), c=dt, factors=[F(v1), F(v0, 2)]),
           dict(test=(1, 0), c=dt, factors=[F(v1), F(v1, 2)])]
    fields = {"prev0": (0, ns_prev0), "prev1": (1, ns_prev1), "old0": (0, ns_old0), "old1": (1, ns_old1)}
    return Problem("ns", golden, ("square", n), ["p2", "p2", "p1"], "qf5pT", True, bil, lin,
                   [dict(comp=0, where="boundary", value=ns_u0_bv),
                    dict(comp=1, where="boundary", value=ns_u1_bv),
                    dict(comp=2, where="pin", value=0.0)], fields)


def laplace2(n, golden=None):
    lin = [dict(test=(0, 0), factors=[("x", 1.0 + x0)]), dict(test=(1, 0), factors=[("x", X.sin(3.0 * x1))])]
    return Problem("laplace2", golden, ("square", n), ["p1", "p1"], "qf5pT", True,
                   lap_terms(2, 0, 0) + lap_terms(2, 1, 1), lin,
                   [dict(comp=0, where="boundary", value=0.0), dict(comp=1, where="boundary", value=0.0)])


GOLDEN_CASES = [
    poisson("p1", 10, "poisson_p1_n10"),
    poisson("p1", 33, "poisson_p1_n33"),
    poisson("p2", 6, "poisson_p2_n6"),
    poisson("p1b", 5, "poisson_p1b_n5"),
    poissong("p1", 8, "poissong_p1_n8"),
    poissong("p2", 5, "poissong_p2_n5"),
    tet("p1", 4, "tet_p1_n4"),
    tet("p1b", 3, "tet_p1b_n3"),
    stokes(6, "stokes_n6"),
    navier_stokes(5, "ns_n5"),
    laplace2(6, "laplace2_n6"),
]


def pin_cell(n):
    """test/stokes_2d_p2_p1.cpp:73: pressure pinned at the first vertex of cell n*n/2 + n/2."""
    return n * n // 2 + n // 2


