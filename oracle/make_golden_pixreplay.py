"""Generates tests/golden/pixreplay.npz by running the REAL reference class
(/root/reference/agent/diffsrdrq/helper_functions/efficient_buffer.py: EfficientReplayBuffer) on a seeded synthetic stream
that wraps the ring twice, and checks the restatement (oracle/pixel_replay_oracle.py) against it bit for bit.
    python -m oracle.make_golden_pixreplay        (build container only: needs /root/reference)"""
import importlib.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.pixel_replay_oracle import OraclePixelReplay, synthetic_stream  # noqa: E402

REF = "/root/reference/agent/diffsrdrq/helper_functions/efficient_buffer.py"


def main():
    spec = importlib.util.spec_from_file_location("ref_efficient_buffer", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    N, B, nstep, gamma, fs = 61, 24, 3, 0.99, 3
    ref = mod.EfficientReplayBuffer(N, B, nstep, gamma, fs)
    ora = OraclePixelReplay(N, B, nstep, gamma, fs)
    stream = synthetic_stream(150, frame_stack=fs, seed=5)
    out = {}
    for t, ts in enumerate(stream):
        ref.add(ts)
        ora.add(ts)
        assert ref.index == ora.index and ref.full == ora.full and np.array_equal(ref.valid, ora.valid), t
        if t in (40, 100, 149):
            np.random.seed(100 + t)
            idx = np.random.choice(ref.valid.nonzero()[0], size=B)
            want = ref.gather_nstep_indices(idx)
            got = ora.gather(idx)
            for a, b in zip(want, got):
                assert a.dtype == b.dtype and np.array_equal(a, b)
            out[f"idx_{t}"] = idx
            for name, a in zip(("obs", "act", "rew", "dis", "nobs", "sobs"), want):
                out[f"{name}_{t}"] = a
            out[f"valid_{t}"] = ref.valid.copy()
            out[f"len_{t}"] = np.int64(len(ref))
    np.savez_compressed(ROOT / "tests" / "golden" / "pixreplay.npz", **out)
    print("wrote tests/golden/pixreplay.npz; restatement bit-identical to the reference at", sorted(k for k in out if k.startswith("idx")))


if __name__ == "__main__":
    main()
