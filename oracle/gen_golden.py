"""TEST INFRASTRUCTURE: regenerate tests/golden/*.npz from the UNMODIFIED reference.

Runs oracle/_ref/tfel_ref (the reference compiled from /root/reference against the
stand-in third-party headers, see oracle/Makefile) on small instances of every problem
the parity tests use and stores what the reference produced. Needs /root/reference, so it
only runs in the build container; the fixtures travel with the repo.

    python oracle/gen_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from refdump import read_dump  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF = os.path.join(HERE, "_ref", "tfel_ref")

# (file stem, problem, fe, n, solver used for the stored reference solution)
CASES = [
    ("poisson_p1_n10", "poisson", "p1", 10, "gmres"),
    ("poisson_p1_n33", "poisson", "p1", 33, "cg"),
    ("poisson_p2_n6", "poisson", "p2", 6, "gmres"),
    ("poisson_p1b_n5", "poisson", "p1b", 5, "gmres"),
    ("poissong_p1_n8", "poissong", "p1", 8, "gmres"),
    ("poissong_p2_n5", "poissong", "p2", 5, "gmres"),
    ("tet_p1_n4", "tet", "p1", 4, "gmres"),
    ("tet_p1b_n3", "tet", "p1b", 3, "gmres"),
    ("stokes_n6", "stokes", "", 6, "gmres"),
    ("ns_n5", "ns", "", 5, "gmres"),
    ("laplace2_n6", "laplace2", "", 6, "gmres"),
]


def run_reference(problem, fe, n, out, solve="", extra=()):
    cmd = [REF, "--problem", problem, "--n", str(n), "--out", out]
    if fe:
        cmd += ["--fe", fe]
    if solve:
        cmd += ["--solve", solve, "--solve-rtol", "1e-13"]
    cmd += list(extra)
    res = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, universal_newlines=True)
    return json.loads(res.stdout.strip().splitlines()[-1])


def main():
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    os.makedirs(GOLDEN, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "tables.bin")
        run_reference("tables", "", 0, path)
        np.savez_compressed(os.path.join(GOLDEN, "ref_tables.npz"), **read_dump(path))
        for stem, problem, fe, n, solve in CASES:
            path = os.path.join(tmp, stem + ".bin")
            meta = run_reference(problem, fe, n, path, solve)
            arrays = read_dump(path)
            arrays["meta_json"] = np.frombuffer(json.dumps(
                {k: v for k, v in meta.items() if not k.startswith("t_") and "per_s" not in k}).encode(), dtype=np.uint8)
            np.savez_compressed(os.path.join(GOLDEN, stem + ".npz"), **arrays)
            print(stem, "nnz", meta.get("nnz"), "its", meta.get("solve_its"), "rnorm", meta.get("solve_rnorm"))


def export_fixtures():
    """tests/golden/export/*: the files the reference's own exporters write (export.hpp) for nodal interpolants on small
    meshes; the case files embed the relative prefix "export/", so the reference runs with tests/golden as its cwd."""
    import shutil
    out = os.path.join(GOLDEN, "export")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    subprocess.check_call([REF, "--problem", "export", "--n", "3", "--out", "export"], cwd=GOLDEN)
    print("export:", len(os.listdir(out)), "files")


if __name__ == "__main__":
    main()
    export_fixtures()
