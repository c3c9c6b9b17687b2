"""Host-side sharding rules of the multi-GPU path (one process per GPU), mirrored from csrc/mvicp.cu so that they
can be tested on CPU with gloo: frame ownership, edge ownership, and the zero-padded all-reduce that acts as an
order-independent all-gather of per-edge blocks."""
import numpy as np


def frame_owner(frame, world, n_frames):
    """owner = frame * world / n_frames (block distribution; csrc/mvicp.cu owner_of)."""
    return (frame * world) // max(1, n_frames)


def owned_edges(edges, rank, world, n_frames, fixed=None):
    """Edges processed by `rank`: those whose src frame it owns and whose src is not fixed (frame.cpp:93)."""
    out = []
    for e, (s, d) in enumerate(edges):
        if fixed is not None and fixed[s]:
            continue
        if frame_owner(s, world, n_frames) == rank:
            out.append(e)
    return out


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
