"""Multi-process path: world_size-2 gloo test of the host logic on CPU; NCCL tests when >= 2 GPUs are visible."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def launch(n, args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)


def test_gloo_world_size_2_host_logic():
    r = launch(2, ["gloo"], 29611)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok") == 2


def test_strip_segments_tile_everything():
    from tfel_b200 import dist as tdist
    offs = [[0, 1089, 4225, 4225, 4225], [0, 1089, 1089, 1089, 1089]]
    for P in (1, 2, 3, 4, 8):
        segs = sorted(s for r in range(P) for s in tdist.strip_segments(offs, P, r))
        at = 0
        for b, e in segs:
            assert b == at
            at = e
        assert at == 4225 + 1089


@pytest.mark.gpu
@pytest.mark.parametrize("which,n", [("poisson_p1", 96), ("poissong_p2", 40), ("tet_p1", 10), ("stokes", 16)])
def test_nccl_distributed_matches_single_gpu(which, n):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    ranks = min(torch.cuda.device_count(), 4)
    r = launch(ranks, ["nccl", which, str(n)], 29621)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-6000:]
