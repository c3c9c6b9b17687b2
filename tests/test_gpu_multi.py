"""Sharded path on real GPUs (skipped unless >= 2 devices): edges split over 2 ranks (balanced by query count), per-edge pair
matrices exchanged through peer memory or ncclAllReduce; poses must be BIT-IDENTICAL to the single-GPU run (DESIGN.md
section 5), and the engine's ownership must be the rule mv_lm_icp_b200/dist.py states."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from helpers import scene
from mv_lm_icp_b200 import synth

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _run(rank, world, port, out_dir, flags):
    import torch.distributed as dist
    import mv_lm_icp_b200 as mv
    from mv_lm_icp_b200.dist import broadcast_unique_id
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sc = scene(6, 20011, 22)
    edges = synth.ring_edges(6, 2)
    eng = mv.Engine(device=rank, flags=flags)
    eng.set_frames(sc["pts"], sc["nor"]); eng.set_graph(edges)
    eng.comm_init(broadcast_unique_id(mv.nccl_unique_id, rank, device="cuda"), rank, world)
    eng.set_poses(sc["poses_init"])
    out = []
    for _ in range(3):
        s = eng.icp_round(0.05, mv.PARAM_SE3, mv.COST_P2PLANE, True)
        out.append((eng.get_poses(), s["num_iterations"]))
    # ownership: get_nn succeeds exactly on the edges dist.edge_owners gives this rank
    from mv_lm_icp_b200.dist import edge_owners
    own = edge_owners(edges, [len(p) for p in sc["pts"]], world, [1] + [0] * 5)
    for e in range(len(edges)):
        try:
            eng.get_nn(e); mine = True
        except RuntimeError:
            mine = False
        assert mine == (own[e] == rank), (e, rank, own[e])
    np.save(os.path.join(out_dir, f"poses_{rank}.npy"), np.stack([o[0] for o in out]))
    np.save(os.path.join(out_dir, f"iters_{rank}.npy"), np.array([o[1] for o in out]))
    eng.close()
    dist.barrier(); dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("flags", [0, 2], ids=["peer-memory", "nccl-only"])
def test_two_gpus_bit_identical_to_one(tmp_path, flags):
    import mv_lm_icp_b200 as mv
    sc = scene(6, 20011, 22)
    edges = synth.ring_edges(6, 2)
    eng = mv.Engine(); eng.set_frames(sc["pts"], sc["nor"]); eng.set_graph(edges); eng.set_poses(sc["poses_init"])
    ref = []
    for _ in range(3):
        s = eng.icp_round(0.05, mv.PARAM_SE3, mv.COST_P2PLANE, True)
        ref.append((eng.get_poses(), s["num_iterations"]))
    eng.close()
    mp.spawn(_run, args=(2, _free_port(), str(tmp_path), flags), nprocs=2, join=True)
    for r in range(2):
        P = np.load(tmp_path / f"poses_{r}.npy"); it = np.load(tmp_path / f"iters_{r}.npy")
        assert it.tolist() == [x[1] for x in ref]
        assert np.array_equal(P.view(np.uint64), np.stack([x[0] for x in ref]).view(np.uint64))
