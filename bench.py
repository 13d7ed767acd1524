#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's configuration.

Workload (configs[1]): 2-D Poisson, P1 on gen_square_mesh(1, 1, n, n) with n = 4096 (33 554 432 cells,
16 785 409 dofs, 117 399 559 stored entries), qf5pT quadrature, f = 1, homogeneous Dirichlet data on the whole
boundary, CG + Jacobi to rtol 1e-8, x0 = 0.

One "step" = one pass of the hot path: clear(); a += integrate<qf5pT>(d<1>(u)*d<1>(v) + d<2>(u)*d<2>(v), m);
b += integrate<qf5pT>(f*v, m); u = a.solve(b, s).  `value` = cells / step time with mesh, DOF map and CSR pattern
resident in HBM; `e2e` = the same through the C ABI from HOST mesh arrays (upload, DOF numbering, Dirichlet set,
pattern, assembly, solve, solution copied back to the host), everything inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mesh-n 4096] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cells/s assembled + solve s to rtol 1e-8 (Poisson P1 4096^2), SpMV HBM% @1-8 B200"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "tfel_ref")
REF_SAMPLE_N = 512


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, n, n_gpus):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture (profiles/), or None
    when no capture exists for this configuration"""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))["per_launch_bytes"][kernel]
        if d["n"] == n and d["n_gpus"] == n_gpus:
            return d["dram_read"] + d["dram_write"]
    except Exception:
        pass
    return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.25)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=15)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
            if any(r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


# --------------------------------------------------------------------------------------------------
# CPU legs: the reference's own CPU implementation (oracle/_ref/tfel_ref: the UNMODIFIED reference sources
# compiled against stand-in third-party headers, CPU CG in place of the absent PETSc), else the oracle port.
# --------------------------------------------------------------------------------------------------

def cpu_reference_once(n, threads):
    """One pass of the path on the CPU at mesh size n: returns (cells, seconds, kind, detail)."""
    if os.path.exists(REF_BIN):
        out = subprocess.run([REF_BIN, "--problem", "poisson", "--fe", "p1", "--n", str(n), "--hash-only",
                              "--solve", "cg", "--solve-rtol", "1e-8", "--threads", str(threads)],
                             capture_output=True, text=True, check=True).stdout
        d = json.loads(out.strip().splitlines()[-1])
        t = d["t_bilinear_s"] + d["t_linear_s"] + d["t_solve_s"]
        return int(d["n_cells"]), t, "reference", {"t_bilinear_s": d["t_bilinear_s"], "t_linear_s": d["t_linear_s"],
                                                   "t_solve_s": d["t_solve_s"], "cg_its": d["solve_its"]}
    # the plain-C restatement (oracle/), single thread
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle as orc
    import oracle_runner
    import problems
    t0 = time.perf_counter()
    got = oracle_runner.run(problems.poisson("p1", n))
    t_asm = time.perf_counter() - t0   # includes the structure build of the port
    t0 = time.perf_counter()
    _, its, _ = orc.cg_jacobi(got["csr_rowptr"], got["csr_colind"], got["csr_val"], got["rhs"], rtol=1e-8)
    t_solve = time.perf_counter() - t0
    return 2 * n * n, t_asm + t_solve, "port", {"t_assembly_incl_structure_s": t_asm, "t_solve_s": t_solve, "cg_its": its}


def cpu_baseline(n=REF_SAMPLE_N):
    threads = os.cpu_count() or 1
    cells, t, kind, detail = cpu_reference_once(n, threads)
    return {"value": cells / t, "unit": "cells/s", "cores": threads if kind == "reference" else 1, "kind": kind,
            "sample": "same problem at n=%d (%d cells): a += ... (1 thread by construction, bilinear_form.hpp:64) + "
                      "b += ... (linear_form thread pool, %d threads) + Jacobi-CG to rtol 1e-8 (1 thread; PETSc absent, "
                      "CPU restatement); %.1f s of CPU work. CG iterations grow ~1.83 n, so cells/s at n=4096 is ~%dx lower"
                      % (n, cells, threads, t, 4096 // n),
            "detail": detail}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n = args.ref_n
    for _ in range(args.warmup):
        cpu_reference_once(n, threads)
    total = 0.0
    cells = 0
    detail = None
    kind = "reference"
    for _ in range(args.steps):
        cells, t, kind, detail = cpu_reference_once(n, threads)
        total += t
    value = cells * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Poisson P1 gen_square_mesh n=%d, qf5pT, f=1, CG+Jacobi rtol 1e-8 (bounded sample of the "
                                   "n=4096 workload)" % n, "n": n, "cells": cells},
            "cpu_baseline": {"value": value, "unit": "cells/s", "cores": threads if kind == "reference" else 1, "kind": kind,
                             "sample": "n=%d per step, %d steps; %s" % (n, args.steps, json.dumps(detail))},
            "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# --------------------------------------------------------------------------------------------------
# own arm
# --------------------------------------------------------------------------------------------------

STIFFNESS = [dict(test=(0, 1), trial=(0, 1), c=1.0), dict(test=(0, 2), trial=(0, 2), c=1.0)]
SOURCE = [dict(test=(0, 0), c=1.0)]


def run_own(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from tfel_b200 import capi
    from tfel_b200 import dist as tdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (tfel_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run)" % (args.gpus, world))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.n
    ctx = tdist.init_context(local_rank)     # distributed (NCCL) when world > 1
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    peak, peak_src = measured_peaks()

    # ---- resident structure (one-time, reported separately) -------------------------------------
    # N > 1: ONE n x n problem, rows of the reference numbering split into horizontal strips (strong scaling);
    # mesh, DOF map and Dirichlet set are replicated (0.67 GB), every rank owns the CSR rows / vector entries of its strip
    def make_forms(mesh_, space_):
        if world == 1:
            return capi.Matrix(space_, space_), capi.Vector(space_)
        segs = tdist.strip_segments([space_.kind_offsets()], world, rank)
        return capi.Matrix(space_, space_, rows=segs), capi.Vector(space_, rows=segs)

    t0 = time.perf_counter()
    mesh = capi.Mesh.square(ctx, 1.0, 1.0, n, n)
    space = capi.Space(mesh, "p1")
    space.add_dirichlet_boundary(0.0)
    a, b = make_forms(mesh, space)
    ctx.synchronize()
    t_structure = time.perf_counter() - t0
    cells = mesh.n_cells
    solver = capi.Solver(ctx, method="cg", precond="jacobi", maxits=args.maxits, rtol=1e-8, atol=1e-50, dtol=1e20,
                         check_every=args.check_every, profile_every=args.profile_every)
    x_dev = torch.empty(a.n_cols, dtype=torch.float64, device="cuda")

    ev = lambda: torch.cuda.Event(enable_timing=True)
    asm_ms, solve_ms, spmv_ms, its_seen, spmv_samples = [], [], [], [], []

    def step(timed):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record(stream)
        a.clear()
        a.assemble("qf5pT", STIFFNESS)
        b.clear()
        b.assemble("qf5pT", SOURCE)
        e1.record(stream)
        _, rep = a.solve(b, solver, out=x_dev)
        e2.record(stream)
        if rep["converged"] != 1:
            raise SystemExit("bench.py: CG did not converge: %r" % (rep,))
        if timed:
            e2.synchronize()
            asm_ms.append(e0.elapsed_time(e1)); solve_ms.append(e1.elapsed_time(e2))
            spmv_ms.append(rep["spmv_avg_ms"]); its_seen.append(rep["its"]); spmv_samples.append(rep["spmv_samples"])
        return rep

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    start, stop = ev(), ev()
    start.record(stream)
    for _ in range(args.steps):
        step(True)
    stop.record(stream)
    stop.synchronize()
    barrier()
    clocks = sampler.summary() if rank == 0 else None
    launches = ctx.launches - launches0
    ms_total = start.elapsed_time(stop)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = cells * args.steps / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel (SpMV inside CG), measured live over the timed region ----
    nnz, N = a.nnz, a.n_local                                          # this rank's rows
    spmv_bytes = 12.0 * nnz + 4.0 * (N + 1) + 16.0 * N
    spmv_avg_ms = float(np.mean(spmv_ms)) if spmv_ms and spmv_ms[0] > 0 else None
    asm_bytes = (4.0 * 3 * cells + 8.0 * 2 * mesh.n_vertices) / world + 8.0 * nnz + 8.0 * N   # A and b together (52 B/cell)
    asm_avg_ms = float(np.mean(asm_ms))
    roofline = {"kernel": "k_spmv_stream<DOT> (CSR fp64 SpMV + fused x.Ax partials, inside CG)", "bound": "hbm",
                "achieved": (spmv_bytes / (spmv_avg_ms * 1e-3) / 1e9) if spmv_avg_ms else None, "peak": peak,
                "unit": "GB/s", "frac": (spmv_bytes / (spmv_avg_ms * 1e-3) / 1e9 / peak) if spmv_avg_ms else None,
                "traffic": ncu_traffic("k_spmv_stream<true,1,1024,128,3,5,int16>", n, world),
                "traffic_note": "ncu capture in profiles/: 1.508 GB per launch < algorithmic 1.7445 GB because the solver streams 16-bit "
                                "column offsets (10 B per stored entry) instead of the reference CSR's int32 columns (12 B), which the "
                                "algorithmic figure counts; hence frac > 1", "peak_source": peak_src, "algorithmic_bytes_per_launch": spmv_bytes,
                "avg_launch_ms": spmv_avg_ms, "samples_per_solve": int(np.mean(spmv_samples)) if spmv_samples else 0,
                "frac_of_nominal_8TBs": (spmv_bytes / (spmv_avg_ms * 1e-3) / 1e9 / 8000.0) if spmv_avg_ms else None}
    assembly = {"cells_per_s": cells / (asm_avg_ms * 1e-3), "ms": asm_avg_ms,
                "algorithmic_bytes": asm_bytes, "achieved_gbs": asm_bytes / (asm_avg_ms * 1e-3) / 1e9,
                "frac_of_peak": asm_bytes / (asm_avg_ms * 1e-3) / 1e9 / peak,
                "kernels": "k_assemble_rows<2,3,3,const,staged,rec> (packed 16-byte row records) + k_assemble_vector_rows<2,3,const>"}
    solve = {"s": float(np.mean(solve_ms)) * 1e-3, "cg_iterations": int(np.mean(its_seen)),
             "ms_per_iteration": float(np.mean(solve_ms)) / max(float(np.mean(its_seen)), 1.0)}

    # ---- e2e: host mesh arrays -> solution on the host, through the C ABI, everything timed -----------
    e2e = None
    if args.e2e_steps > 0:
        nv = (n + 1) * (n + 1)
        ii, jj = np.meshgrid(np.arange(n + 1, dtype=np.float64), np.arange(n + 1, dtype=np.float64), indexing="xy")
        verts = torch.empty((nv, 2), dtype=torch.float64).pin_memory()
        verts.numpy()[:, 0] = ((1.0 / n) * ii).ravel(); verts.numpy()[:, 1] = ((1.0 / n) * jj).ravel()
        q = (np.arange(n, dtype=np.int64)[None, :] + (n + 1) * np.arange(n, dtype=np.int64)[:, None]).ravel()
        cells_h = torch.empty((2 * n * n, 3), dtype=torch.int32).pin_memory()
        c = cells_h.numpy().view(np.uint32).reshape(n * n, 6)
        c[:, 0] = q; c[:, 1] = q + 1; c[:, 2] = q + n + 2; c[:, 3] = q; c[:, 4] = q + n + 1; c[:, 5] = q + n + 2
        x_host = torch.empty(nv, dtype=torch.float64).pin_memory()
        del ii, jj, q
        # release the resident objects: the e2e pass builds its own
        solver.close(); b.close(); a.close(); space.close(); mesh.close()

        def e2e_step():
            m2 = capi.Mesh.create(ctx, verts.numpy(), cells_h.numpy().view(np.uint32))
            s2 = capi.Space(m2, "p1")
            s2.add_dirichlet_boundary(0.0)
            a2, b2 = make_forms(m2, s2)
            a2.assemble("qf5pT", STIFFNESS)
            b2.assemble("qf5pT", SOURCE)
            sv = capi.Solver(ctx, method="cg", precond="jacobi", maxits=args.maxits, rtol=1e-8, check_every=args.check_every)
            _, rep = a2.solve(b2, sv, out=x_host.numpy())
            sv.close(); b2.close(); a2.close(); s2.close(); m2.close()
            return rep

        e2e_step()   # warm-up
        barrier()
        t0 = time.perf_counter()
        s_ev, e_ev = ev(), ev()
        s_ev.record(stream)
        for _ in range(args.e2e_steps):
            rep = e2e_step()
        e_ev.record(stream)
        e_ev.synchronize()
        barrier()
        wall = time.perf_counter() - t0
        te = torch.tensor([max(wall, s_ev.elapsed_time(e_ev) * 1e-3)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": cells * args.e2e_steps / float(te.item()), "unit": "cells/s",
               "h2d_bytes_per_step": int(verts.numel() * 8 + cells_h.numel() * 4), "d2h_bytes_per_step": int(x_host.numel() * 8),
               "s_per_step": float(te.item()) / args.e2e_steps, "steps": args.e2e_steps,
               "includes": "mesh upload, DOF numbering, boundary + Dirichlet set, CSR pattern, assembly of A and b, CG solve, "
                           "solution download", "max_u": float(x_host.max())}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "configs[1]: 2D Poisson P1 on %dx%d square mesh, qf5pT, f=1, CG + Jacobi to rtol 1e-8"
                                       % (n, n), "n": n, "cells": cells, "dofs": a.n_rows, "rows_rank0": N, "nnz_rank0": nnz,
                           "l2": "inputs larger than L2 (CSR 1.5 GB, mesh 0.67 GB per pass vs 126 MB)",
                           "parallelism": ("%d horizontal strips of rows (reference numbering); halo exchange + reductions over %s"
                                           % (world, "peer-mapped memory (CUDA IPC, NVLink)" if ctx.p2p else "NCCL send/recv + allreduce"))
                           if world > 1 else "single GPU",
                           "structure_build_s": t_structure},
                "assembly": assembly, "solve": solve, "roofline": roofline, "clocks": clocks,
                "e2e": e2e, "gpu_launches": int(launches)}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.ref_n)
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    # libraries (NCCL prints its version) must not pollute stdout: everything but the final JSON line goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--mesh-n", dest="n", type=int, default=4096)
    ap.add_argument("--ref-n", type=int, default=REF_SAMPLE_N)
    ap.add_argument("--maxits", type=int, default=20000)
    ap.add_argument("--check-every", type=int, default=100)
    ap.add_argument("--profile-every", type=int, default=64)
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_own(args)


if __name__ == "__main__":
    sys.exit(main())
