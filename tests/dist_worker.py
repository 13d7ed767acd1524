"""Worker of the multi-process tests (one process per rank; launched by torch.distributed.run).

mode "gloo": CPU only -- NCCL id exchange through torch.distributed and the row partition (host logic).
mode "nccl": one GPU per rank -- distributed Poisson assemble + CG against the single-GPU result of rank 0.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

from tfel_b200 import capi, dist as tdist  # noqa: E402


def main():
    mode = sys.argv[1]
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", "0"))
    if mode == "gloo":
        dist.init_process_group("gloo")
        nccl_id = tdist.exchange_nccl_id(capi.Context.nccl_unique_id)
        ids = [None] * world
        dist.all_gather_object(ids, nccl_id)
        assert len(nccl_id) == 128 and all(i == ids[0] for i in ids)
        # the partition every rank computes tiles the rows
        offs = [[0, 25, 81, 81, 81], [0, 25, 81, 81, 81], [0, 25, 25, 25, 25]]   # P2, P2, P1 on a 4x4 square
        mine = tdist.strip_segments(offs, world, rank)
        allseg = [None] * world
        dist.all_gather_object(allseg, mine)
        flat = sorted(s for segs in allseg for s in segs if s[1] > s[0])
        at = 0
        for b, e in flat:
            assert b == at, (flat,)
            at = e
        assert at == 81 + 81 + 25
        assert allseg == tdist.all_segments(offs, world)
        dist.destroy_process_group()
        print("gloo worker %d/%d ok" % (rank, world))
        return

    import cuda_runner
    import problems
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = tdist.init_context(local)
    which = sys.argv[2] if len(sys.argv) > 2 else "poisson_p1"
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 96
    problem = {"poisson_p1": problems.poisson("p1", n), "poisson_p2": problems.poisson("p2", n),
               "poissong_p2": problems.poissong("p2", n), "tet_p1": problems.tet("p1", n),
               "stokes": problems.stokes(n)}[which]
    # every rank builds mesh + spaces (replicated structure), owns a strip of rows
    B = cuda_runner.build(ctx, problem)          # full matrix (reference for this test) ...
    cuda_runner.assemble(B)
    rp, ci, val = B.matrix.csr()
    b_full = B.vector.download()
    segs = tdist.strip_segments([s.kind_offsets() for s in B.spaces], world, rank)
    A = capi.Matrix(B.spaces, B.spaces, rows=segs)   # ... and the distributed one
    b = capi.Vector(B.spaces, rows=segs)
    A.assemble(problem.quad, cuda_runner.lower(problem.bilinear), B.fields)
    b.assemble(problem.quad, cuda_runner.lower(problem.linear), B.fields)
    lrp, lci, lval = A.csr()
    rows = np.concatenate([np.arange(s[0], s[1]) for s in segs])
    assert A.n_local == rows.shape[0]
    # the owned rows are exactly the rows of the full reference CSR (bit-exact structure AND values: same kernels,
    # same per-row order)
    for i in (0, len(rows) // 2, len(rows) - 1):
        g = rows[i]
        assert np.array_equal(lci[lrp[i]:lrp[i + 1]], ci[rp[g]:rp[g + 1]])
    lens = np.diff(rp)[rows]
    assert np.array_equal(np.diff(lrp), lens)
    sel = np.concatenate([np.arange(rp[g], rp[g + 1]) for g in rows]) if rows.size < 200000 else None
    if sel is not None:
        assert np.array_equal(lci, ci[sel]) and np.array_equal(lval, val[sel])
    assert np.array_equal(b.download(), b_full[rows])
    if which == "stokes":
        kw = dict(method="cg", precond="jacobi", maxits=50, rtol=1e-30)   # not SPD: compare a fixed number of SpMVs only
    else:
        kw = dict(method="cg", precond="jacobi", maxits=20000, rtol=1e-10, check_every=10)
    # distributed SpMV against the single-GPU one
    s1 = capi.Solver(ctx, **kw); s1.set_operator(B.matrix)
    sd = capi.Solver(ctx, **kw); sd.set_operator(A)
    xg = np.random.RandomState(5).standard_normal(B.matrix.n_rows)
    y_full = s1.spmv(xg)
    y_loc = sd.spmv(xg[rows])
    assert np.abs(y_loc - y_full[rows]).max() <= 1e-13 * np.abs(y_full).max()
    assert np.array_equal(sd.gather(xg[rows]), xg)
    if which != "stokes":
        x_full, rep1 = B.matrix.solve(B.vector, s1)
        x_dist, repd = A.solve(b, sd)
        assert rep1["converged"] == 1 and repd["converged"] == 1, (rep1, repd)
        assert abs(int(rep1["its"]) - int(repd["its"])) <= 2, (rep1["its"], repd["its"])
        err = np.linalg.norm(x_dist - x_full) / np.linalg.norm(x_full)
        assert err <= 1e-8, err
        if rank == 0:
            print("dist worker: %s n=%d ranks=%d its=%d/%d rel diff %.2e" % (which, n, world, repd["its"], rep1["its"], err))
    # restarted GMRES on the row-distributed operator (halo exchange before every product, rank-ordered reductions of
    # every MGS dot product) against the single-GPU GMRES: same Arnoldi process up to the rounding of the partial sums
    for pc in ("jacobi", "block_jacobi", "ilu0"):   # ilu0: ILU(0) of every rank's diagonal block
        if which == "stokes":
            gkw = dict(method="gmres", precond=pc, maxits=45, restart=20, rtol=1e-30)   # a fixed number of Arnoldi steps
        else:
            gkw = dict(method="gmres", precond=pc, maxits=4000, restart=60, rtol=1e-10)
        if pc == "block_jacobi":
            gkw["block_size"] = 4
        g1 = capi.Solver(ctx, **gkw); g1.set_operator(B.matrix)
        gd = capi.Solver(ctx, **gkw); gd.set_operator(A)
        xg_full, r1 = B.matrix.solve(B.vector, g1)
        xg_dist, rd = A.solve(b, gd)
        if pc == "jacobi":   # block-Jacobi / ILU(0) factorise rank-local blocks: another preconditioner when rows are split
            assert abs(int(r1["its"]) - int(rd["its"])) <= 2, (r1["its"], rd["its"])
        if which != "stokes":
            assert r1["converged"] == 1 and rd["converged"] == 1, (r1, rd)
            gerr = np.linalg.norm(xg_dist - xg_full) / np.linalg.norm(xg_full)
            assert gerr <= 1e-8, gerr
        elif pc == "jacobi":
            gerr = np.linalg.norm(xg_dist - xg_full) / max(np.linalg.norm(xg_full), 1e-300)
            assert gerr <= 1e-6, gerr
        if rank == 0:
            print("dist worker: %s n=%d ranks=%d gmres+%s its=%d/%d" % (which, n, world, pc, rd["its"], r1["its"]))
        g1.close(); gd.close()
    s1.close(); sd.close(); A.close(); b.close(); cuda_runner.close(B)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
