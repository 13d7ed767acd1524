"""Row-parallel symbolic ILU(k) -- host prototype of the device kernel planned for the next round (tools / tests only).

Level-of-fill ILU with the sum rule keeps entry (i, j) iff G(A) has a path i -> ... -> j of at most k + 1 edges whose
intermediate vertices are all smaller than min(i, j) (the incomplete fill-path theorem). That turns the inherently
sequential row-merge (oracle/tfel_oracle.c: ilu_build; PETSc's PCFactorSetLevels, solver.hpp:162-163) into one independent
search per row:

  * U part of row i (j > i): breadth-first search from i, expanding only through vertices t < i (every intermediate of a
    path to some j > i has to be below min(i, j) = i), at most k + 1 edges deep; every vertex j > i met on the way is an entry.
  * L part (j < i): the intermediates have to be below j, which depends on the target -- but (i, j) with j < i is an entry
    iff (j, i) is in the U part of row j of A^T, so the same search runs on the transposed pattern and its result is
    transposed back.

One warp per row on the device (frontier and visited set in shared memory), a radix sort of the emitted (row << 32 | col)
keys, and the numeric factorisation / level-scheduled solves that already run on any CSR pattern (krylov.cu: Ilu0).
"""
import numpy as np


def transpose_pattern(n, rowptr, colind):
    counts = np.bincount(colind, minlength=n)
    t_rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=t_rowptr[1:])
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    order = np.lexsort((rows, colind))
    return t_rowptr, rows[order].astype(np.int64)


def upper_fill(n, rowptr, colind, levels):
    """for every row i the set {j > i : (i, j) is an entry of ILU(levels)} -- independent per row"""
    out = []
    for i in range(n):
        depth = {i: 0}
        frontier = [i]
        found = []
        for d in range(1, levels + 2):
            nxt = []
            for h in frontier:
                for t in colind[rowptr[h]:rowptr[h + 1]]:
                    t = int(t)
                    if t in depth:
                        continue
                    depth[t] = d
                    if t > i:
                        found.append(t)
                    elif d < levels + 1:
                        nxt.append(t)        # t < i may be an intermediate vertex of a longer path
            frontier = nxt
        out.append(sorted(found))
    return out


def iluk_pattern(rowptr, colind, levels):
    """(rowptr, colind) of ILU(levels), columns ascending, diagonal always present"""
    rowptr = np.asarray(rowptr, dtype=np.int64); colind = np.asarray(colind, dtype=np.int64)
    n = rowptr.shape[0] - 1
    upper = upper_fill(n, rowptr, colind, levels)
    t_rowptr, t_colind = transpose_pattern(n, rowptr, colind)
    lower_t = upper_fill(n, t_rowptr, t_colind, levels)      # row j of A^T: the i > j with (i, j) an entry
    rows = [[i] for i in range(n)]
    for i in range(n):
        rows[i].extend(upper[i])
    for j in range(n):
        for i in lower_t[j]:
            rows[i].append(j)
    out_rp = np.zeros(n + 1, dtype=np.int32)
    out_ci = []
    for i in range(n):
        r = sorted(set(rows[i]))
        out_ci.extend(r)
        out_rp[i + 1] = len(out_ci)
    return out_rp, np.asarray(out_ci, dtype=np.int32)
