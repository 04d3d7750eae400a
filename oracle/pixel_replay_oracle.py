"""CPU restatement of the reference's pixel replay ring with its n-step frame-stack gather.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Follows agent/diffsrdrq/helper_functions/efficient_buffer.py
(`EfficientReplayBuffer`): `_initial_setup` :53-64, `add_data_point` :66-105, `__next__` :111-114,
`gather_nstep_indices` :116-143, `__len__` :145-149.  Plain numpy on the host, written as functions over one state object.
Pinned against the real reference class by oracle/make_golden_pixreplay.py -> tests/golden/pixreplay.npz
(tests/test_oracle_golden.py re-checks it wherever the tests run).

One deliberate difference: when a trajectory's first frame ends exactly at the last slot the reference leaves
`index == buffer_size` and raises IndexError on the next add (:75-93 never wrap that case); the restatement wraps to 0 like
every other path does.  The fixture does not exercise that corner.
"""
from __future__ import annotations

import collections

import numpy as np

TimeStep = collections.namedtuple("TimeStep", ["is_first", "observation", "action", "reward", "discount"])
TimeStep.first = lambda self: self.is_first  # the reference calls time_step.first()


class OraclePixelReplay:
    def __init__(self, buffer_size, batch_size, nstep, discount, frame_stack):  # efficient_buffer.py:36-51
        self.buffer_size, self.batch_size, self.nstep, self.discount = buffer_size, batch_size, nstep, discount
        self.frame_stack = frame_stack
        self.index, self.traj_index, self.full = -1, 0, False
        self.discount_vec = np.power(discount, np.arange(nstep)).astype("float32")
        self.next_dis = discount ** nstep

    def _initial_setup(self, ts):  # :53-64
        self.index = 0
        self.obs_shape = list(ts.observation.shape)
        self.ims_channels = self.obs_shape[0] // self.frame_stack
        self.act_shape = ts.action.shape
        self.obs = np.zeros([self.buffer_size, self.ims_channels, *self.obs_shape[1:]], dtype=np.uint8)
        self.act = np.zeros([self.buffer_size, *self.act_shape], dtype=np.float32)
        self.rew = np.zeros([self.buffer_size], dtype=np.float32)
        self.dis = np.zeros([self.buffer_size], dtype=np.float32)
        self.valid = np.zeros([self.buffer_size], dtype=np.bool_)

    def add(self, ts):  # :66-109
        if self.index == -1:
            self._initial_setup(ts)
        N, fs = self.buffer_size, self.frame_stack
        latest = ts.observation[-self.ims_channels:]
        if ts.first():
            end_index = self.index + fs
            end_invalid = end_index + fs + 1
            for k in range(fs):  # frame_stack copies of the first frame, wrapping
                self.obs[(self.index + k) % N] = latest
            if end_invalid > N:
                if end_index > N:
                    end_index %= N
                    self.full = True
                end_invalid %= N
                self.valid[self.index:N] = False
                self.valid[0:end_invalid] = False
            else:
                self.valid[self.index:end_invalid] = False
            if end_index == N:
                end_index, self.full = 0, True
            self.index, self.traj_index = end_index, 1
        else:
            self.obs[self.index] = latest
            self.act[self.index] = ts.action
            self.rew[self.index] = ts.reward
            self.dis[self.index] = ts.discount
            self.valid[(self.index + fs) % N] = False
            if self.traj_index >= self.nstep:
                self.valid[(self.index - self.nstep + 1) % N] = True
            self.index += 1
            self.traj_index += 1
            if self.index == N:
                self.index, self.full = 0, True

    def sample_indices(self):  # :113
        return np.random.choice(self.valid.nonzero()[0], size=self.batch_size)

    def gather(self, indices):  # :116-143
        N, fs, n = self.buffer_size, self.frame_stack, self.nstep
        rng = np.stack([np.arange(i - fs, i + n) for i in indices], axis=0) % N
        steps, obs_r, nobs_r, sobs_r = rng[:, fs:], rng[:, :fs], rng[:, -fs:], rng[:, 1:fs + 1]
        rew = np.sum(self.rew[steps] * self.discount_vec, axis=1, keepdims=True)
        shape = [len(indices), *self.obs_shape]
        obs, nobs, sobs = (np.reshape(self.obs[r], shape) for r in (obs_r, nobs_r, sobs_r))
        act = self.act[indices]
        dis = np.expand_dims(self.next_dis * self.dis[nobs_r[:, -1]], axis=-1)
        return obs, act, rew, dis, nobs, sobs

    def __len__(self):  # :145-149
        return self.buffer_size if self.full else self.index


def synthetic_stream(n_steps, frame_stack=3, c=3, hw=12, action_dim=4, seed=0, episode_len=17):
    """Deterministic stream of dm_env-like time steps: episodes of `episode_len` steps, uint8 frame stacks whose newest frame
    is fresh noise (older frames are the previous newest ones, as a frame-stack wrapper produces them)."""
    rng = np.random.default_rng(seed)
    out, stack = [], None
    for t in range(n_steps):
        first = t % episode_len == 0
        frame = rng.integers(0, 256, size=(c, hw, hw), dtype=np.uint8)
        stack = np.concatenate([frame] * frame_stack) if first else np.concatenate([stack[c:], frame])
        out.append(TimeStep(first, stack.copy(), rng.uniform(-1, 1, action_dim).astype(np.float32),
                            np.float32(rng.standard_normal()), np.float32(1.0 if (t + 1) % episode_len else 0.0)))
    return out
