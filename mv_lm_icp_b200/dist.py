"""Host-side sharding rules of the multi-GPU path (one process per GPU), mirrored from csrc/mvicp.cu so that they
can be tested on CPU with gloo: edge ownership (balanced by query count), and the zero-padded all-reduce that acts as an
order-independent all-gather of per-edge blocks."""
import numpy as np


def edge_owners(edges, n_pts, world, fixed=None):
    """Rank that processes each edge (-1: src frame fixed, nobody) -- csrc/mvicp.cu assign_edge_owners: the edges whose
    src is not fixed, in graph order, are cut into `world` contiguous runs of (nearly) equal query count; an edge goes to
    the rank that the midpoint of its query range falls to."""
    act = [e for e, (s, d) in enumerate(edges) if fixed is None or not fixed[s]]
    total = sum(int(n_pts[edges[e][0]]) for e in act)
    out = [-1] * len(edges)
    before = 0
    for e in act:
        n = int(n_pts[edges[e][0]])
        out[e] = min(world - 1, ((2 * before + n) * world) // (2 * total)) if total > 0 else 0
        before += n
    return out


def owned_edges(edges, rank, world, n_pts, fixed=None):
    """Edges processed by `rank` (frame.cpp:93: a fixed frame computes no correspondences)."""
    return [e for e, o in enumerate(edge_owners(edges, n_pts, world, fixed)) if o == rank]


def broadcast_unique_id(make_id, rank, device=None):
    """Rank 0 creates the 128-byte NCCL id (mvicp_nccl_unique_id), torch.distributed ships it to every rank."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(make_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


def gather_blocks_by_allreduce(local_blocks, owned, n_edges):
    """What the LM loop does between lm_reduce_kernel and lm_step_kernel: every rank contributes its own edges' blocks,
    zeros elsewhere, and a SUM all-reduce returns all blocks.  x + 0 is exact, so the result is bit-identical on every
    rank and independent of the number of ranks."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros((n_edges, local_blocks.shape[1]), dtype=torch.float64)
    buf[owned] = torch.from_numpy(np.ascontiguousarray(local_blocks))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf.numpy()
