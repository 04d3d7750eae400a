"""Drop-in for the reference's `utils.buffer.ReplayBuffer` (reference: utils/buffer.py:13-48), backed by a
device-resident fp32 ring (`rlrep_ring_*` in include/rlrep_b200.h).

Same surface: `ReplayBuffer(state_dim, action_dim, max_size)`, `.add(s, a, s', r, done)`,
`.sample(batch_size) -> Batch(state, action, reward, next_state, done)` with fp32 tensors on the device, and the
public counters `size`, `ptr`, `max_size`.  Indices are drawn from the global legacy numpy RNG exactly like the
reference (`np.random.randint(0, size, size=batch_size)`), so seeding numpy identically yields the same batches.

`add` stages rows in pinned-free host memory and ships them in batches (row-at-a-time H2D would dominate once the
ring lives in HBM); `sample` / `agent.train` flush first, so the observable behaviour is unchanged.
"""
from __future__ import annotations

import collections
import ctypes as C

import numpy as np

from . import _lib

Batch = collections.namedtuple("Batch", ["state", "action", "reward", "next_state", "done"])  # buffer.py:7-10

_STAGE_ROWS = 1024


class ReplayBuffer:
    def __init__(self, state_dim, action_dim, max_size=int(1e6)):
        import torch
        if not torch.cuda.is_available():
            raise _lib.RlrepError("rlrep_b200.ReplayBuffer needs a CUDA device (there is no CPU fallback)")
        self._lib = _lib.load()
        self.max_size = int(max_size)
        self.state_dim, self.action_dim = int(state_dim), int(action_dim)
        self.ptr = 0
        self.size = 0
        self.device = torch.device("cuda", torch.cuda.current_device())
        h = C.c_void_p()
        _lib.check(self._lib.rlrep_ring_create(self.state_dim, self.action_dim, self.max_size, C.byref(h)))
        self._h = h
        vals = [C.c_int() for _ in range(5)]
        _lib.check(self._lib.rlrep_ring_layout(self._h, *[C.byref(v) for v in vals]))
        self.record_floats, self.off_action, self.off_reward, self.off_done, self.off_next_state = (v.value for v in vals)
        self._stage = np.zeros((_STAGE_ROWS, self.record_floats), dtype=np.float32)
        self._pending = 0

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and getattr(self, "_lib", None) is not None:
            self._lib.rlrep_ring_destroy(h)

    # -- reference surface -------------------------------------------------------------------------------------
    def add(self, state, action, next_state, reward, done):
        row = self._stage[self._pending]
        S, A = self.state_dim, self.action_dim
        row[:S] = state
        row[self.off_action:self.off_action + A] = action
        row[self.off_reward] = reward
        row[self.off_done] = done
        row[self.off_next_state:self.off_next_state + S] = next_state
        self._pending += 1
        self.ptr = (self.ptr + 1) % self.max_size
        self.size = min(self.size + 1, self.max_size)
        if self._pending == _STAGE_ROWS:
            self.flush()

    def sample(self, batch_size) -> Batch:
        ind = np.random.randint(0, self.size, size=batch_size)
        return self.take(ind)

    # -- extras ------------------------------------------------------------------------------------------------
    def flush(self):
        """Ship staged rows to the device ring."""
        if self._pending:
            _lib.check(self._lib.rlrep_ring_add_packed(self._h, self._stage.ctypes.data, self._pending, None))
            self._pending = 0

    def take(self, ind) -> Batch:
        """Gather the given rows (bit-exact fp32 casts of what was added)."""
        import torch
        self.flush()
        ind = np.ascontiguousarray(ind, dtype=np.int64)
        out = torch.empty((len(ind), self.record_floats), dtype=torch.float32, device=self.device)
        _lib.check(self._lib.rlrep_ring_gather(self._h, ind.ctypes.data, len(ind), out.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream))
        S, A = self.state_dim, self.action_dim
        return Batch(state=out[:, :S], action=out[:, self.off_action:self.off_action + A],
                     reward=out[:, self.off_reward:self.off_reward + 1],
                     next_state=out[:, self.off_next_state:self.off_next_state + S],
                     done=out[:, self.off_done:self.off_done + 1])

    def load(self, state, action, next_state, reward, done):
        """Bulk fill from the reference's five column arrays (float64 or float32), starting at slot 0."""
        self._pending = 0
        n = len(state)
        f64 = np.asarray(state).dtype == np.float64
        dt = np.float64 if f64 else np.float32
        cols = [np.ascontiguousarray(a, dtype=dt) for a in (state, action, next_state, reward, done)]
        _lib.check(self._lib.rlrep_ring_load(self._h, *[c.ctypes.data for c in cols], n, int(f64), None))
        self.size = min(n, self.max_size)
        self.ptr = n % self.max_size
