"""One process per GPU: rendez-vous of the NCCL communicator through torch.distributed and the row partition.

The path shards by contiguous blocks of the reference DOF numbering: vertex dofs are numbered row-major
(`mesh.cpp:19-51`: id = j (n + 1) + i) and edge dofs by ascending (low vertex, high vertex), so an equal split of every
entity kind of every component is a horizontal strip of the square mesh (a k-slab of the cube mesh). Pure host logic:
no CUDA needed to compute a partition (tested on CPU with the gloo backend).
"""
import os


def split(count, n_ranks, rank):
    """[begin, end) of `rank` in an equal split of range(count)"""
    return (count * rank) // n_ranks, (count * (rank + 1)) // n_ranks


def strip_segments(kind_offsets_per_component, n_ranks, rank):
    """Global row segments rank owns: an equal share of every entity kind of every component.

    kind_offsets_per_component: for each component the 5 offsets of Space.kind_offsets(); components are stacked in the
    block layout [u0 | u1 | p] (composite_bilinear_form.hpp:41-62)."""
    segs = []
    base = 0
    for offs in kind_offsets_per_component:
        for sd in range(4):
            cnt = offs[sd + 1] - offs[sd]
            if cnt <= 0:
                continue
            b, e = split(cnt, n_ranks, rank)
            segs.append((base + offs[sd] + b, base + offs[sd] + e))
        base += offs[4]
    return segs


def all_segments(kind_offsets_per_component, n_ranks):
    return [strip_segments(kind_offsets_per_component, n_ranks, r) for r in range(n_ranks)]


def exchange_nccl_id(make_id, backend_group=None):
    """Rank 0 creates the 128-byte NCCL unique id; everybody receives it through torch.distributed (any backend)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(make_id()), dtype=torch.uint8))
    dist.broadcast(buf, src=0, group=backend_group)
    return bytes(buf.cpu().numpy().tobytes())


def init_context(local_rank=None):
    """tfel_b200 context of this process: distributed when torch.distributed is initialised with world_size > 1."""
    import torch.distributed as dist
    from . import capi
    if local_rank is None:
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return capi.Context(local_rank)
    nccl_id = exchange_nccl_id(capi.Context.nccl_unique_id)
    return capi.Context(local_rank, rank=dist.get_rank(), n_ranks=dist.get_world_size(), nccl_id=nccl_id)
