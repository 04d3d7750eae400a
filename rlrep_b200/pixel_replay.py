"""Drop-in for the reference's `EfficientReplayBuffer` (agent/diffsrdrq/helper_functions/efficient_buffer.py:35-149) with the
frames resident in HBM (`rlrep_pixring_*` in include/rlrep_b200.h).

Same constructor, `add(time_step)`, `__next__()` and `__len__()`.  `time_step` is duck-typed like the reference's dm_env
wrapper: `.first()`, `.observation` (uint8 [frame_stack * c, H, W]), `.action`, `.reward`, `.discount`.  The bookkeeping of
which slots may be sampled (`valid`, `index`, `traj_index`, `full`) is the reference's, line for line in behaviour; indices
are drawn with `np.random.choice(valid.nonzero()[0], size=batch_size)` from the global numpy RNG like the reference does.
`__next__` returns the reference's tuple (obs, act, rew, dis, nobs, sobs) as CUDA tensors assembled by one gather kernel:
the pixel agents' update calls take them as they are, so no frame crosses PCIe between the replay and the update.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class PixelReplayBuffer:
    def __init__(self, buffer_size, batch_size, nstep, discount, frame_stack, data_specs=None):
        self.buffer_size = int(buffer_size)
        self.index = -1
        self.traj_index = 0
        self.frame_stack = int(frame_stack)
        self._recorded_frames = self.frame_stack + 1
        self.batch_size = int(batch_size)
        self.nstep = int(nstep)
        self.discount = discount
        self.full = False
        self.discount_vec = np.power(discount, np.arange(nstep)).astype("float32")  # efficient_buffer.py:50
        self.next_dis = discount ** nstep
        self._lib, self._h = None, None

    # -- efficient_buffer.py:53-64
    def _initial_setup(self, time_step):
        import torch
        if not torch.cuda.is_available():
            raise _lib.RlrepError("rlrep_b200.PixelReplayBuffer needs a CUDA device (there is no CPU fallback)")
        self.index = 0
        self.obs_shape = list(time_step.observation.shape)
        self.ims_channels = self.obs_shape[0] // self.frame_stack
        self.act_shape = np.asarray(time_step.action).shape
        self._frame_shape = (self.ims_channels, *self.obs_shape[1:])
        self._frame_bytes = int(np.prod(self._frame_shape))
        self._A = int(np.prod(self.act_shape)) if self.act_shape else 1
        self.valid = np.zeros([self.buffer_size], dtype=np.bool_)
        self._lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device())
        h = C.c_void_p()
        _lib.check(self._lib.rlrep_pixring_create(self.buffer_size, self._frame_bytes, self._A, self.frame_stack, self.nstep,
                                                  C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and getattr(self, "_lib", None) is not None:
            self._lib.rlrep_pixring_destroy(h)

    def _write(self, slot, copies, frame, action=None, reward=0.0, discount=0.0):
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        act = None if action is None else np.ascontiguousarray(action, dtype=np.float32).reshape(-1)
        _lib.check(self._lib.rlrep_pixring_write(self._h, int(slot), int(copies), frame.ctypes.data,
                                                 act.ctypes.data if act is not None else None, float(reward),
                                                 float(discount), int(act is not None)))

    # -- efficient_buffer.py:66-105: identical index / valid arithmetic; only the frame stores go to the device ring
    def add_data_point(self, time_step):
        first = time_step.first()
        latest_obs = time_step.observation[-self.ims_channels:]
        if first:
            end_index = self.index + self.frame_stack
            end_invalid = end_index + self.frame_stack + 1
            self._write(self.index, self.frame_stack, latest_obs)  # frame_stack copies, wrapping modulo the ring
            if end_invalid > self.buffer_size:
                if end_index > self.buffer_size:
                    end_index = end_index % self.buffer_size
                    self.full = True
                end_invalid = end_invalid % self.buffer_size
                self.valid[self.index:self.buffer_size] = False
                self.valid[0:end_invalid] = False
            else:
                self.valid[self.index:end_invalid] = False
            if end_index == self.buffer_size:  # the reference leaves index == buffer_size here and fails on the next add
                end_index, self.full = 0, True
            self.index = end_index
            self.traj_index = 1
        else:
            self._write(self.index, 1, latest_obs, time_step.action, time_step.reward, time_step.discount)
            self.valid[(self.index + self.frame_stack) % self.buffer_size] = False
            if self.traj_index >= self.nstep:
                self.valid[(self.index - self.nstep + 1) % self.buffer_size] = True
            self.index += 1
            self.traj_index += 1
            if self.index == self.buffer_size:
                self.index = 0
                self.full = True

    def add(self, time_step):
        if self.index == -1:
            self._initial_setup(time_step)
        self.add_data_point(time_step)

    def __next__(self):
        indices = np.random.choice(self.valid.nonzero()[0], size=self.batch_size)  # efficient_buffer.py:113
        return self.gather_nstep_indices(indices)

    def __iter__(self):
        return self

    def gather_nstep_indices(self, indices):
        """efficient_buffer.py:116-143 on the device: (obs, act, rew, dis, nobs, sobs) as CUDA tensors."""
        import torch
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        n = idx.shape[0]
        u8 = lambda: torch.empty((n, *self.obs_shape), dtype=torch.uint8, device=self.device)
        obs, nobs, sobs = u8(), u8(), u8()
        act = torch.empty((n, *self.act_shape), dtype=torch.float32, device=self.device)
        rew = torch.empty((n, 1), dtype=torch.float32, device=self.device)
        dis = torch.empty((n, 1), dtype=torch.float32, device=self.device)
        torch.cuda.current_stream().synchronize()  # the outputs were just allocated on torch's stream
        dv = np.ascontiguousarray(self.discount_vec, dtype=np.float32)
        _lib.check(self._lib.rlrep_pixring_gather(self._h, idx.ctypes.data, n, dv.ctypes.data, float(np.float32(self.next_dis)),
                                                  obs.data_ptr(), act.data_ptr(), rew.data_ptr(), dis.data_ptr(),
                                                  nobs.data_ptr(), sobs.data_ptr()))
        return obs, act, rew, dis, nobs, sobs

    def __len__(self):
        return self.buffer_size if self.full else self.index
