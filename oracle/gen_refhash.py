"""TEST INFRASTRUCTURE: full-size fingerprints of what the UNMODIFIED reference produces.

Runs oracle/_ref/tfel_ref --hash-only (no array dump) on instances too large for array fixtures and stores its JSON line:
position-weighted 64-bit hashes of every integer array (dof_map, g2l, Dirichlet dofs, CSR rowptr / colind) and
sum / sum of squares / sum of absolute values of every floating-point array, plus the reference's own phase timings
on this container's CPU (bilinear assembly is single-threaded by construction, bilinear_form.hpp:64).

    python oracle/gen_refhash.py            # hours: add_dirichlet_boundary is O(#boundary x N) (fes.hpp:112-113);
                                            # poisson n=4096 took 35 min here, 1799 s of it in add_dirichlet_boundary

The hash (ref_driver.cpp dumper::hash_int):  H = sum_i (a_i + 1) (2 i + 1) 0x9E3779B97F4A7C15  (mod 2^64).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref", "tfel_ref")
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = [("poisson_p1_n1024", ["--problem", "poisson", "--fe", "p1", "--n", "1024"]),
         ("poisson_p1_n2048", ["--problem", "poisson", "--fe", "p1", "--n", "2048"]),
         ("poisson_p1_n4096", ["--problem", "poisson", "--fe", "p1", "--n", "4096"]),
         ("stokes_n256", ["--problem", "stokes", "--n", "256"]),
         ("tet_p1_n48", ["--problem", "tet", "--fe", "p1", "--n", "48"]),
         # a1 / a2 / a3 / a4 alone (vertices, cells, geometry, face neighbours, boundary submesh): minutes, not hours
         ("mesh_square_n1024", ["--problem", "mesh", "--n", "1024"]),
         ("mesh_square_n4096", ["--problem", "mesh", "--n", "4096"]),
         ("mesh_cube_n48", ["--problem", "meshcube", "--n", "48"]),
         ("mesh_cube_n142", ["--problem", "meshcube", "--n", "142"])]

if __name__ == "__main__":
    for name, args in CASES:
        out = subprocess.run([REF] + args + ["--hash-only", "--threads", "8"], capture_output=True, text=True, check=True).stdout
        with open(os.path.join(GOLDEN, "refhash_%s.json" % name), "w") as f:
            f.write(out.strip().splitlines()[-1] + "\n")
        print(name, "done")
