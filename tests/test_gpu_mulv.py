"""GPU: the muLV-Rep DrQ-v2 pixel update (agent/mulvdrq/drqv2.py:313-461) through the C ABI against the CPU oracle
(oracle/mulv_oracle.py, pinned to the real reference class by tests/golden/mulvdrq_b4.npz) on the same weights, batches
and seeds: forward metrics, the gradients of every parameter tensor (read back as "grad/<name>"), parameters and
Polyak targets after the update, and a second update through the `up_every` protocol."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# channels, action_dim, feat_dim, hid_dim, batch; the last one is mulv_config.py's full size (b_size 256, hid_dim 1024)
CONFIGS = [(9, 4, 100, 64, 4), (9, 4, 100, 128, 32), (3, 6, 50, 64, 32), (9, 4, 100, 1024, 256)]
GROUPS = ("encoder", "predict_encoder", "decoder", "feat_encoder", "feat_decoder", "feat_f", "critic", "actor")
FORWARD = ("critic_loss", "critic_q1", "critic_q2", "critic_target_q", "s_loss", "r_loss", "kl_loss")
# fp32: the bars are the DrQ-v2 test's (tests/test_gpu_drq.py) -- tensors behind ReLU stacks see a handful of mask flips;
# actor gradients are taken after the other groups' Adam step and inherit its +-lr element flips.
# tf32: operand rounding puts ~1e-3 of the ReLU units on the other side of zero; each flip changes a gradient by a term.
# The critic's gradients are sums over rows of sign-mixed TD errors that largely cancel, so a flipped unit weighs
# ~1/sqrt(B) of a column instead of 1/B (measured 3e-2 .. 1.2e-1; the same kernels give 4.5e-6 with fp32 operands).
# Parameters after the update are compared norm-wise: Adam's first step is lr * sign(g) for every element, so elements whose
# gradient is within rounding distance of zero (inputs that are almost always behind a ReLU) move by +-lr on either side
# (measured 7.7e-4 on actor.trunk.0.weight [100, 39200] at the full size, 1e-4 elsewhere).
TOL = {"fp32": dict(fwd=1e-5, grad=5e-3, critic=5e-3, actor=5e-3, param=2e-3, after=2e-2),
       "tf32": dict(fwd=1e-2, grad=1e-1, critic=2.5e-1, actor=1.5e-1, param=2e-2, after=1e-1)}


def _cfg(F, H):
    return dict(aug=True, pre_aug=False, back_q2feat=True, tanh=True, both_q=False, q_activ="relu", q_loss="huber",
                q_up_n=1, l2_norm=0.0, c_targ_tau=0.01, up_every=2, stddev_schedule="linear(1.0,0.1,500000)",
                stddev_clip=0.3, feat_dim=F, hid_dim=H, lr=1e-4, vae_w=0.5, mse_w=1.0, c_noise=0.1)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


@pytest.mark.parametrize("C,A,F,H,B", CONFIGS)
@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_mulv_update_matches_oracle(C, A, F, H, B, precision):
    from oracle import mulv_oracle as M
    from rlrep_b200.pixel import MuLVDrQv2
    tol = dict(TOL[precision])
    if precision == "tf32" and B >= 256:
        tol["fwd"] = 1e-3  # the north_star TF32 bar on the losses at mulv_config.py's own size; at B <= 32 the Q means are
        #                    ~1e-2 in magnitude and a few TF32 roundings of single rows show at 2e-3
    init = M.init_state(C, A, F, H, seed=0)
    oracle = M.OracleMuLVDrQ(A, init)
    agent = MuLVDrQv2((C, 84, 84), (A,), _cfg(F, H), precision=precision)
    agent.load_state_dict(init)
    batches = [M.synthetic_pixel_batch(B, C, 84, A, seed=30 + i) for i in range(2)]

    # ---- first update from identical weights
    torch.manual_seed(1)
    o = oracle.update(batches[0], step=0)
    rng_after_oracle = torch.get_rng_state()
    torch.manual_seed(1)
    c = agent.update(iter([tuple(batches[0])]), step=0)
    assert torch.equal(torch.get_rng_state(), rng_after_oracle), "the shim must consume the generator like the reference"
    fwd = {k: max(0.0, abs(c[k] - o[k]) - 1e-5) / (abs(o[k]) + 1e-12) for k in FORWARD}
    act = max(0.0, abs(c["actor_loss"] - o["actor_loss"]) - 1e-5) / (abs(o["actor_loss"]) + 1e-12)
    g, ref = agent.grads(), oracle.last_grads
    assert set(ref) <= set(g), sorted(set(ref) - set(g))
    per_tensor = {k: _rel(g[k], r) for k, r in ref.items()}
    worst = {grp: max((e, k) for k, e in per_tensor.items() if k.split(".")[0] == grp) for grp in GROUPS}
    csd, osd = agent.state_dict(), oracle.state_dict()
    assert set(osd) <= set(csd), sorted(set(osd) - set(csd))
    perr = {k: _rel(csd[k], v) for k, v in osd.items()}
    pworst = max((e, k) for k, e in perr.items())
    print(f"\nmulvdrq C={C} A={A} F={F} H={H} B={B} {precision}: forward " + ", ".join(f"{k} {v:.1e}" for k, v in fwd.items())
          + f"; actor_loss {act:.1e}\n  grads: " + ", ".join(f"{grp} {e:.1e} ({k})" for grp, (e, k) in worst.items())
          + f"\n  worst param after update {pworst[0]:.1e} ({pworst[1]}); {agent.gpu_launches_last_update} launches/update")
    bad = sorted(((e, k) for k, e in per_tensor.items()), reverse=True)[:8]
    print("  largest gradient errors: " + ", ".join(f"{k} {e:.1e}" for e, k in bad))
    assert max(fwd.values()) < tol["fwd"], fwd
    assert act < tol["after"], act
    for grp, (e, k) in worst.items():
        assert e < tol.get(grp, tol["grad"]), (grp, k, e)
    assert pworst[0] < tol["param"], pworst

    # ---- up_every = 2: step 1 is a no-op that draws nothing, step 2 updates again
    assert agent.update(iter([tuple(batches[1])]), step=1) == {} and oracle.update(batches[1], step=1) == {}
    torch.manual_seed(2)
    o2 = oracle.update(batches[1], step=2)
    torch.manual_seed(2)
    c2 = agent.update(iter([tuple(batches[1])]), step=2)
    after = {k: max(0.0, abs(c2[k] - o2[k]) - 1e-5) / (abs(o2[k]) + 1e-12) for k in FORWARD + ("actor_loss",)}
    csd, osd = agent.state_dict(), oracle.state_dict()
    pworst2 = max((_rel(csd[k], v), k) for k, v in osd.items())
    print(f"  second update: metrics {max(after.values()):.1e}; worst param {pworst2[0]:.1e} ({pworst2[1]})")
    assert max(after.values()) < tol["after"], after
    assert pworst2[0] < 2 * tol["param"], pworst2
    agent.close()


def test_mulv_act_matches_oracle():
    from oracle import mulv_oracle as M
    from rlrep_b200.pixel import MuLVDrQv2
    C, A, F, H, B = 9, 4, 100, 64, 4
    init = M.init_state(C, A, F, H, seed=0)
    oracle = M.OracleMuLVDrQ(A, init)
    agent = MuLVDrQv2((C, 84, 84), (A,), _cfg(F, H), precision="fp32")
    agent.load_state_dict(init)
    agent.prepare(B)
    obs = M.synthetic_pixel_batch(1, C, 84, A, seed=3).img[0]
    assert np.allclose(agent.act(obs, 1000, True), oracle.act(obs, 1000, True), atol=2e-5)
    for step in (250000, 100):  # past / inside the exploration phase (uniform actions replace the sample)
        torch.manual_seed(4)
        a_c = agent.act(obs, step, False)
        torch.manual_seed(4)
        a_o = oracle.act(obs, step, False)
        assert np.allclose(a_c, a_o, atol=2e-5), (step, a_c, a_o)
    agent.close()
