"""GPU: the device-resident pixel replay ring (rlrep_b200.PixelReplayBuffer over rlrep_pixring_*) against the restated
reference ring (oracle/pixel_replay_oracle.py, bit-identical to the real EfficientReplayBuffer): same adds, same sampled
indices -> the six gathered tensors, the valid mask and the length are bit-exact; and a pixel agent's update on a batch
that never left the device equals its update on the same batch handed over from the host."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,hw,fs,nstep", [(61, 12, 3, 3), (257, 84, 3, 3), (100, 10, 2, 1), (90, 7, 4, 5)])
def test_ring_and_gather_are_bit_exact(N, hw, fs, nstep):
    from oracle.pixel_replay_oracle import OraclePixelReplay, synthetic_stream
    from rlrep_b200 import PixelReplayBuffer
    B = 40
    ora = OraclePixelReplay(N, B, nstep, 0.99, fs)
    buf = PixelReplayBuffer(N, B, nstep, 0.99, fs)
    checks = 0
    for t, ts in enumerate(synthetic_stream(int(2.6 * N), frame_stack=fs, hw=hw, seed=N, episode_len=23)):
        ora.add(ts)
        buf.add(ts)
        assert buf.index == ora.index and buf.full == ora.full and len(buf) == len(ora)
        if t % 37 == 36 and ora.valid.any():
            assert np.array_equal(buf.valid, ora.valid)
            np.random.seed(t)
            idx = ora.sample_indices()
            want = ora.gather(idx)
            np.random.seed(t)
            got = next(buf)
            for name, a, b in zip(("obs", "act", "rew", "dis", "nobs", "sobs"), want, got):
                assert b.is_cuda and tuple(b.shape) == a.shape, (name, b.shape, a.shape)
                assert torch.equal(b.cpu(), torch.from_numpy(a)), (t, name)
            checks += 1
    assert checks >= 3


def test_update_from_device_batch_equals_update_from_host_batch():
    """Frames gathered on the device feed MuLVDrQv2.update / DrQv2.train_step by device pointer (no PCIe round trip): the
    metrics and parameters must equal those of the same update fed from host arrays."""
    from oracle import mulv_oracle as M
    from oracle.pixel_replay_oracle import synthetic_stream
    from rlrep_b200 import PixelReplayBuffer
    from rlrep_b200.pixel import MuLVDrQv2
    C, A, F, H, B = 9, 4, 50, 64, 16
    cfg = dict(aug=True, pre_aug=False, back_q2feat=True, tanh=True, both_q=False, q_activ="relu", q_loss="huber", q_up_n=1,
               l2_norm=0.0, c_targ_tau=0.01, up_every=1, stddev_schedule="linear(1.0,0.1,500000)", stddev_clip=0.3,
               feat_dim=F, hid_dim=H, lr=1e-4, vae_w=0.5, mse_w=1.0, c_noise=0.1)
    buf = PixelReplayBuffer(300, B, 3, 0.99, 3)
    for ts in synthetic_stream(260, frame_stack=3, hw=84, action_dim=A, seed=2, episode_len=50):
        buf.add(ts)
    np.random.seed(0)
    dev_batch = next(buf)
    host_batch = tuple(t.cpu().numpy() for t in dev_batch)
    outs = []
    for batch in (dev_batch, host_batch):
        agent = MuLVDrQv2((C, 84, 84), (A,), cfg, precision="fp32")
        agent.load_state_dict(M.init_state(C, A, F, H, seed=0))
        torch.manual_seed(3)
        m = agent.update(iter([batch]), step=0)
        outs.append((m, agent.state_dict()))
        agent.close()
    assert outs[0][0] == outs[1][0]
    for k, v in outs[1][1].items():
        assert torch.equal(outs[0][1][k], v), k
