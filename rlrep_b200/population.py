"""Populations of independent agents on one GPU (BASELINE.json configs[4]: 64 seeds = 8 agents per GPU x 8 GPUs).

The reference trains one agent per process (`main.py`, `agent/mulvdrq/train_metaworld.py`); a seed sweep is N processes.
Here N agent handles live in ONE process per GPU: every handle owns its stream, weights, optimiser state and staging
buffers (`include/rlrep_b200.h`: "one handle = one agent = one stream"), nothing is shared and nothing is communicated.
`Population.map(fn)` runs `fn(i, agent_i)` for all agents at once from one host thread per agent -- the C entry points
release the GIL and pin themselves to the handle's device -- so the GPU co-schedules kernels of different agents and the
launch gaps of one agent's update are filled by the others'.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor


class Population:
    def __init__(self, agents):
        self.agents = list(agents)
        if not self.agents:
            raise ValueError("a population needs at least one agent")
        self._pool = ThreadPoolExecutor(max_workers=len(self.agents), thread_name_prefix="rlrep-agent")

    def __len__(self):
        return len(self.agents)

    def map(self, fn):
        """fn(i, agent) for every member concurrently; returns the results in member order (exceptions propagate)."""
        futures = [self._pool.submit(fn, i, a) for i, a in enumerate(self.agents)]
        return [f.result() for f in futures]

    def update_all(self, batches, step):
        """One update of every member: batches[i] is member i's replay batch.  Pixel agents expose `update` (muLV-Rep) or
        `train_step` (DrQ-v2 family); state agents `train(buffer, batch_size)` take (buffer, batch_size) tuples."""
        def one(i, a):
            if hasattr(a, "update"):
                return a.update(iter([batches[i]]), step)
            if hasattr(a, "train_step"):
                return a.train_step(iter([batches[i]]), step)
            buf, bs = batches[i]
            return a.train(buf, bs)
        return self.map(one)

    def close(self):
        self._pool.shutdown(wait=True)
        for a in self.agents:
            if hasattr(a, "close"):
                a.close()
        self.agents = []
