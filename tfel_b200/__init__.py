"""tfel_b200: B200 (sm_100a) assembly + solve path behind TFEL's bilinear_form / linear_form / solve
surface. The product is libtfel_cuda.so (tfel_b200/csrc, C ABI in include/tfel_cuda.h); this package
is its Python binding. Importing it never touches oracle/ and never falls back to the CPU."""
from . import capi  # noqa: F401
from .capi import (Context, Mesh, Space, Matrix, Vector, Solver, integrate, TfelCudaError,  # noqa: F401
                   LIB_PATH, SYMBOLS)
