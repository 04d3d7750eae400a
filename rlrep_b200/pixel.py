"""Host-side mirror of the reference's pixel encoder over librlrep_b200.so's `rlrep_conv_encoder_*` entry points.

`ConvEncoder` stands in for `Encoder` + `RandomShiftsAug` of agent/diffsrdrq/network_arch/drqv2.py:21-57,138-167 (the
same stack is agent/mulvdrq/drqv2.py:19-96): uint8 frame stacks in, flattened 32x35x35 features out, and the backward
pass that autograd would run.  It is the first building block of the pixel agents (kernel K15 of SURVEY.md); the agents
themselves are not assembled yet.  No fallback: without the CUDA library / a device the constructor raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class ConvEncoder:
    LAYERS = ("convnet.0", "convnet.2", "convnet.4", "convnet.6")

    def __init__(self, obs_shape=(9, 84, 84), batch=256, precision="tf32"):
        if not torch.cuda.is_available():
            raise _lib.RlrepError("rlrep_b200.ConvEncoder needs a CUDA device (there is no CPU fallback)")
        c, h, w = obs_shape
        if h != w:
            raise ValueError("square frames only (RandomShiftsAug asserts h == w, drqv2.py:30)")
        self.lib = _lib.load()
        self.obs_shape, self.batch = (int(c), int(h), int(w)), int(batch)
        self._stream = torch.cuda.Stream()
        self._h = C.c_void_p()
        _lib.check(self.lib.rlrep_conv_encoder_create(self.batch, int(c), int(h), _lib.PRECISION[precision],
                                                      self._stream.cuda_stream, C.byref(self._h)))
        d = C.c_int()
        _lib.check(self.lib.rlrep_conv_encoder_feature_dim(self._h, C.byref(d)))
        self.repr_dim = d.value

    def close(self):
        h, self._h = self._h, None
        if h:
            self.lib.rlrep_conv_encoder_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _shape(self, layer, bias):
        cin = self.obs_shape[0] if layer == 0 else 32
        return (32,) if bias else (32, cin, 3, 3)

    def _read(self, layer, what):
        out = np.empty(self._shape(layer, what in (1, 3)), dtype=np.float32)
        _lib.check(self.lib.rlrep_conv_encoder_read(self._h, layer, what, out.ctypes.data))
        return torch.from_numpy(out)

    def state_dict(self):
        sd = {}
        for i, name in enumerate(self.LAYERS):
            sd[name + ".weight"], sd[name + ".bias"] = self._read(i, 0), self._read(i, 1)
        return sd

    def grads(self):
        g = {}
        for i, name in enumerate(self.LAYERS):
            g[name + ".weight"], g[name + ".bias"] = self._read(i, 2), self._read(i, 3)
        return g

    def load_state_dict(self, sd):
        for i, name in enumerate(self.LAYERS):
            for what, key in ((0, ".weight"), (1, ".bias")):
                arr = np.ascontiguousarray(torch.as_tensor(sd[name + key]).detach().cpu().float().numpy())
                if arr.shape != self._shape(i, what == 1):
                    raise ValueError(f"{name + key}: expected {self._shape(i, what == 1)}, got {arr.shape}")
                _lib.check(self.lib.rlrep_conv_encoder_write(self._h, i, what, arr.ctypes.data))

    def forward(self, obs: torch.Tensor, shifts: torch.Tensor | None = None) -> torch.Tensor:
        """obs uint8 [B, C, H, W] on the GPU; shifts int32 [B, 2] = (x, y) in [0, 8] as drawn by RandomShiftsAug's
        `torch.randint(0, 2 * pad + 1, (n, 1, 1, 2))` (drqv2.py:43-47), or None for no augmentation."""
        assert obs.is_cuda and obs.dtype == torch.uint8 and tuple(obs.shape) == (self.batch, *self.obs_shape)
        obs = obs.contiguous()
        if shifts is not None:
            shifts = shifts.to(device=obs.device, dtype=torch.int32).reshape(self.batch, 2).contiguous()
        feat = torch.empty(self.batch, self.repr_dim, device=obs.device, dtype=torch.float32)
        cur = torch.cuda.current_stream()
        self._stream.wait_stream(cur)
        _lib.check(self.lib.rlrep_conv_encoder_forward(self._h, obs.data_ptr(),
                                                       shifts.data_ptr() if shifts is not None else None, feat.data_ptr()))
        cur.wait_stream(self._stream)
        self._keep = (obs, shifts)
        return feat

    def backward(self, dfeat: torch.Tensor) -> None:
        assert dfeat.is_cuda and dfeat.dtype == torch.float32 and tuple(dfeat.shape) == (self.batch, self.repr_dim)
        dfeat = dfeat.contiguous()
        cur = torch.cuda.current_stream()
        self._stream.wait_stream(cur)
        _lib.check(self.lib.rlrep_conv_encoder_backward(self._h, dfeat.data_ptr()))
        cur.wait_stream(self._stream)
        self._keep_d = dfeat
