"""world_size-2 gloo test of the host logic of the sharded path (ownership, id broadcast, block exchange)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mv_lm_icp_b200 import synth
from mv_lm_icp_b200.dist import broadcast_unique_id, frame_owner, gather_blocks_by_allreduce, owned_edges


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    M = 7
    edges = synth.ring_edges(M, 2)
    fixed = [1] + [0] * (M - 1)
    mine = owned_edges(edges, rank, world, M, fixed)
    uid = broadcast_unique_id(lambda: bytes(range(128)), rank)
    assert uid == bytes(range(128))
    rng = np.random.default_rng(1234)                      # same stream on every rank = the "true" per-edge blocks
    truth = rng.normal(size=(len(edges), 56)) * 10.0 ** rng.integers(-8, 8, size=(len(edges), 1))
    got = gather_blocks_by_allreduce(truth[mine], mine, len(edges))
    free = [e for e, (s, d) in enumerate(edges) if not fixed[s]]
    assert np.array_equal(got[free].view(np.uint64), truth[free].view(np.uint64))   # bit-exact gather
    assert not got[[e for e in range(len(edges)) if e not in free]].any()
    np.save(os.path.join(out_dir, f"owned_{rank}.npy"), np.array(mine))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    M = 7
    edges = synth.ring_edges(M, 2)
    a, b = (set(np.load(tmp_path / f"owned_{r}.npy").tolist()) for r in range(2))
    assert not (a & b)
    assert a | b == {e for e, (s, d) in enumerate(edges) if s != 0}
    assert all(frame_owner(edges[e][0], 2, M) == 0 for e in a) and all(frame_owner(edges[e][0], 2, M) == 1 for e in b)


def test_owner_is_a_balanced_block_distribution():
    for M, G in [(20, 1), (20, 2), (20, 4), (20, 8), (40, 4), (64, 8), (5, 8)]:
        own = [frame_owner(f, G, M) for f in range(M)]
        assert own == sorted(own) and max(own) <= G - 1
        counts = np.bincount(own, minlength=G)
        assert counts.max() - counts[counts > 0].min() <= 1 or M < G
