"""GPU, 2+ devices (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu`): the batch-sharded CTRL-SAC
update (NCCL all-gather of mu, reduce-scatter of its gradient, all-reduce of parameter gradients) against the CPU oracle
on the GLOBAL batch, and bit-identical parameters across ranks.  Skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, precision, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        from pathlib import Path
        sys.path.insert(0, str(Path(__file__).resolve().parent))
        from oracle import rl_oracle as O
        from parity_util import Space, worst_info_error
        from rlrep_b200 import ReplayBuffer
        from rlrep_b200.agents import ShardedCTRLSACAgent
        S, A, B, rows, n = 17, 6, 64 * world, 4000, 3
        kw = dict(hidden_dim=128, feature_dim=256, extra_feature_steps=2)
        init = O.init_state("ctrlsac", S, A, kw, seed=0)
        oring = O.synthetic_ring(S, A, rows, seed=0)
        agent = ShardedCTRLSACAgent(S, A, Space(A), discount=0.99, tau=0.005, precision=precision, **kw)
        agent.load_state_dict(init)
        buf = ReplayBuffer(S, A, max_size=rows)
        buf.load(oring.state, oring.action, oring.next_state, oring.reward, oring.done)
        np.random.seed(1)
        torch.manual_seed(1)
        infos = [agent.train(buf, B) for _ in range(n)]
        sd = agent.state_dict()
        digest = {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items()}
        gathered = [None] * world
        dist.all_gather_object(gathered, (infos, digest))
        msg = "ok"
        if rank == 0:
            for r in range(1, world):
                assert gathered[r][0] == infos, f"rank {r} reports different metrics"
                assert gathered[r][1] == digest, f"rank {r} holds different parameters"
            oracle = O.ORACLES["ctrlsac"](S, A, init, discount=0.99, tau=0.005, as_written=False, **kw)
            np.random.seed(1)
            torch.manual_seed(1)
            oi = [oracle.train(oring, B) for _ in range(n)]
            wi, where = worst_info_error(infos, oi, atol=1e-5)
            osd = oracle.state_dict()
            wp, wname = 0.0, None
            for k, v in osd.items():
                if k == "log_alpha":
                    assert abs(float(sd[k]) - float(v)) < 1e-4
                    continue
                d = ((sd[k].double() - v.double()).norm() / (v.double().norm() + 1e-30)).item()
                if d > wp:
                    wp, wname = d, k
            tol = 1e-5 if precision == "fp32" else 1e-3  # north_star bars
            msg = f"ok world={world} {precision}: worst info rel {wi:.2e} at {where}; worst param rel-l2 {wp:.2e} at {wname}"
            assert wi < tol and wp < tol, msg
        agent.close()
        out.put((rank, msg))
    except Exception as e:
        import traceback
        out.put((rank, "FAIL " + repr(e) + traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_sharded_ctrlsac_matches_global_batch_oracle(precision):
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    if world not in (2, 4, 8):
        world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, precision, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(out.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    print(results[0][1])
    assert all(m.startswith("ok") for _, m in results), results
