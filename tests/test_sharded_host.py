"""CPU, world_size 2 over gloo: host-side logic of the batch-sharded CTRL-SAC agent -- the ranks' slices of the global
index / noise draws tile the reference's draws exactly, every rank leaves the global RNGs in the reference's state, and
the NCCL unique id travels through torch.distributed.  (The collectives themselves need GPUs: tests/test_gpu_sharded.py.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class Space:
    def __init__(self, A):
        self.low, self.high = -np.ones(A, np.float32), np.ones(A, np.float32)


class FakeBuffer:
    size = 5000


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rlrep_b200.agents import CTRLSACAgent, ShardedCTRLSACAgent
        kw = dict(hidden_dim=32, feature_dim=64, extra_feature_steps=2)
        B, S, A = 64, 17, 6
        agent = ShardedCTRLSACAgent(S, A, Space(A), **kw)
        np.random.seed(5)
        torch.manual_seed(5)
        idx, eps = agent._draw(FakeBuffer(), B)
        state = (np.random.get_state()[1].copy(), torch.get_rng_state().clone())
        ref = CTRLSACAgent(S, A, Space(A), **kw)
        np.random.seed(5)
        torch.manual_seed(5)
        gidx, geps = ref._draw(FakeBuffer(), B)
        assert np.array_equal(np.random.get_state()[1], state[0]) and torch.equal(torch.get_rng_state(), state[1])
        gathered = [None] * world
        dist.all_gather_object(gathered, (idx, eps))
        K, b = 3, B // world
        full_idx = np.concatenate([g[0].reshape(K, b) for g in gathered], axis=1).reshape(-1)
        full_eps = np.concatenate([g[1].reshape(2, b, A) for g in gathered], axis=1).reshape(-1)
        assert np.array_equal(full_idx, gidx) and np.array_equal(full_eps, geps)
        with pytest.raises(ValueError):
            agent.local_batch(63)
        cfg = agent._config(agent.local_batch(B))
        assert cfg.batch_size == b and cfg.feature_steps == K
        # the id broadcast used by _comm_handle (payload only: creating the communicator needs GPUs)
        box = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        assert box[0] == bytes(range(128))
        out.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_draws_tile_the_reference_draws():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
