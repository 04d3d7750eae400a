"""CPU restatement of the muLV-Rep DrQ-v2 pixel update (reference: agent/mulvdrq/drqv2.py:313-461 `DrQV2Agent.update`
and :284-311 `update_actor`; networks :19-196, agent/mulvdrq/vae.py:13-124) -- TEST INFRASTRUCTURE ONLY.

Groundwork for SURVEY.md 8a row a16: the CUDA path for this agent is not built yet; this file and its fixtures
(tests/golden/mulvdrq_*.npz, generated from the real reference class by oracle/make_golden_mulv.py under the hydra /
omegaconf / matplotlib import shims of SURVEY.md 8c) pin the arithmetic the next round has to reproduce.

Configuration path restated: `mulv_config.py` defaults -- aug=True, pre_aug=False, back_q2feat=True, tanh=True,
both_q=False, q_activ='relu', q_loss='huber', use_feature_target=True, q_up_n=1, c_targ_tau<1 (soft updates), l2_norm=0,
pretrain=False.  RNG order on torch's CPU default generator: shift draw (img), shift draw (next_img),
randn[B, F] (feat_encoder.sample), _standard_normal[B, A] (next action), randn[20, F] (critic_target), randn[20, F]
(critic), _standard_normal[B, A] (actor), randn[20, F] (critic inside the actor step).
"""
from __future__ import annotations

import collections
import math

import numpy as np
import torch
import torch.nn.functional as F

from .drq_oracle import PixelBatch, aug, draw_shift, schedule, synthetic_pixel_batch  # noqa: F401  (shared pieces)

LOG_SIG_MAX, LOG_SIG_MIN = 2, -20  # vae.py:9-10
REPR = 32 * 35 * 35


def layer_table(C, A, feat_dim=100, hid_dim=1024):
    t = []
    for mod, cin in (("encoder", C), ("predict_encoder", 3)):
        for i, ci in zip((0, 2, 4, 6), (cin, 32, 32, 32)):
            t += [(f"{mod}.convnet.{i}.weight", (32, ci, 3, 3)), (f"{mod}.convnet.{i}.bias", (32,))]
    for i in (0, 2, 4, 6):  # ConvTranspose2d weights are [in, out, kh, kw]
        t += [(f"decoder.deconvnet.{i}.weight", (32, 32, 3, 3)), (f"decoder.deconvnet.{i}.bias", (32,))]
    t += [("decoder.deconvnet.8.weight", (3, 32, 2, 2)), ("decoder.deconvnet.8.bias", (3,))]
    t += [("actor.trunk.0.weight", (feat_dim, REPR)), ("actor.trunk.0.bias", (feat_dim,)),
          ("actor.trunk.1.weight", (feat_dim,)), ("actor.trunk.1.bias", (feat_dim,)),
          ("actor.policy.0.weight", (hid_dim, feat_dim)), ("actor.policy.0.bias", (hid_dim,)),
          ("actor.policy.2.weight", (hid_dim, hid_dim)), ("actor.policy.2.bias", (hid_dim,)),
          ("actor.policy.4.weight", (A, hid_dim)), ("actor.policy.4.bias", (A,))]
    for l, (o, i) in zip(("l1", "l2", "l3", "l4", "l5", "l6"),
                         ((hid_dim, feat_dim), (hid_dim, hid_dim), (1, hid_dim)) * 2):
        t += [(f"critic.{l}.weight", (o, i)), (f"critic.{l}.bias", (o,))]
    for mod, inp in (("feat_encoder", 2 * REPR + A), ("feat_f", REPR + A)):
        for head in ("mean_linear", "log_std_linear"):
            t += [(f"{mod}.{head}.0.weight", (feat_dim, inp)), (f"{mod}.{head}.0.bias", (feat_dim,)),
                  (f"{mod}.{head}.1.weight", (feat_dim,)), (f"{mod}.{head}.1.bias", (feat_dim,))]
    t += [("feat_decoder.l1.weight", (hid_dim, feat_dim)), ("feat_decoder.l1.bias", (hid_dim,)),
          ("feat_decoder.l2.weight", (hid_dim, hid_dim)), ("feat_decoder.l2.bias", (hid_dim,)),
          ("feat_decoder.state_linear.weight", (REPR, hid_dim)), ("feat_decoder.state_linear.bias", (REPR,)),
          ("feat_decoder.reward_linear.weight", (1, hid_dim)), ("feat_decoder.reward_linear.bias", (1,))]
    return t


def init_state(C, A, feat_dim=100, hid_dim=1024, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = collections.OrderedDict()
    for name, shape in layer_table(C, A, feat_dim, hid_dim):
        if name.endswith(".1.weight") and len(shape) == 1:      # LayerNorm gain
            sd[name] = 1.0 + 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
        elif name.endswith(".1.bias") and "_linear" in name or name.endswith("trunk.1.bias"):
            sd[name] = 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else int(np.prod(sd[name[:-4] + "weight"].shape[1:]))
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
    return sd


def conv_encoder(p, pre, obs):  # Encoder / PredictEncoder.forward, drqv2.py:69-73, 92-96
    h = obs / 255.0 - 0.5
    h = F.relu(F.conv2d(h, p[pre + ".convnet.0.weight"], p[pre + ".convnet.0.bias"], stride=2))
    for i in (2, 4, 6):
        h = F.relu(F.conv2d(h, p[f"{pre}.convnet.{i}.weight"], p[f"{pre}.convnet.{i}.bias"], stride=1))
    return h.reshape(h.shape[0], -1)


def decoder(p, x):  # Decoder.forward, drqv2.py:105-117
    h = x.view(x.shape[0], 32, 35, 35)
    for i, stride in ((0, 1), (2, 1), (4, 1), (6, 2)):
        h = F.relu(F.conv_transpose2d(h, p[f"decoder.deconvnet.{i}.weight"], p[f"decoder.deconvnet.{i}.bias"], stride=stride))
    return F.conv2d(h, p["decoder.deconvnet.8.weight"], p["decoder.deconvnet.8.bias"], stride=1, padding=1)


def ln_head(p, pre, x, tanh):
    h = F.linear(x, p[pre + ".0.weight"], p[pre + ".0.bias"])
    h = F.layer_norm(h, (h.shape[-1],), p[pre + ".1.weight"], p[pre + ".1.bias"], 1e-5)
    return torch.tanh(h) if tanh else h


def gaussian(p, pre, x):  # vae.Encoder.forward / GaussianFeature.forward (vae.py:40-48, 118-124), tanh=True
    mean = ln_head(p, pre + ".mean_linear", x, True)
    log_std = torch.clamp(ln_head(p, pre + ".log_std_linear", x, False), min=LOG_SIG_MIN, max=LOG_SIG_MAX)
    return mean, log_std


def noisy_critic(p, pre, mean, log_std, c_noise, num_noise=20):  # Critic.forward, drqv2.py:177-196 (relu)
    B, d = mean.shape
    x = mean[:, None, :] + log_std.exp()[:, None, :] * torch.randn([num_noise, d]) * c_noise
    x = x.reshape(-1, d)
    out = []
    for a, b, c in (("l1", "l2", "l3"), ("l4", "l5", "l6")):
        q = F.relu(F.linear(x, p[f"{pre}.{a}.weight"], p[f"{pre}.{a}.bias"])).reshape(B, num_noise, -1).mean(dim=1)
        q = F.relu(F.linear(q, p[f"{pre}.{b}.weight"], p[f"{pre}.{b}.bias"]))
        out.append(F.linear(q, p[f"{pre}.{c}.weight"], p[f"{pre}.{c}.bias"]))
    return out


def actor_dist(p, obs):
    h = ln_head(p, "actor.trunk", obs, True)
    h = F.relu(F.linear(h, p["actor.policy.0.weight"], p["actor.policy.0.bias"]))
    h = F.relu(F.linear(h, p["actor.policy.2.weight"], p["actor.policy.2.bias"]))
    return torch.tanh(F.linear(h, p["actor.policy.4.weight"], p["actor.policy.4.bias"]))


def trunc_sample(mu, std, clip):  # agent_utils.TruncatedNormal.sample, agent_utils.py:117-126
    eps = torch.normal(torch.zeros(mu.shape), torch.ones(mu.shape)) * std
    if clip is not None:
        eps = torch.clamp(eps, -clip, clip)
    x = mu + eps
    return x - x.detach() + torch.clamp(x, -1.0 + 1e-6, 1.0 - 1e-6).detach()


TARGETS = {"critic_target.": "critic.", "encoder_target.": "encoder.", "feat_f_target.": "feat_f."}
OPT_GROUPS = ("encoder", "decoder", "actor", "critic", "predict_encoder", "feat_encoder", "feat_decoder", "feat_f")


class OracleMuLVDrQ:
    def __init__(self, action_dim, state, *, lr=1e-4, c_targ_tau=0.01, up_every=2, vae_w=0.5, mse_w=1.0, c_noise=0.1,
                 stddev_schedule="linear(1.0,0.1,500000)", stddev_clip=0.3):
        self.A, self.tau, self.up_every = action_dim, c_targ_tau, up_every
        self.vae_w, self.mse_w, self.c_noise, self.clip = vae_w, mse_w, c_noise, stddev_clip
        self.sched = schedule(stddev_schedule)
        self.p = {k: v.clone().requires_grad_() for k, v in state.items()}
        self.tgt = {t + k[len(s):]: v.detach().clone() for t, s in TARGETS.items() for k, v in self.p.items()
                    if k.startswith(s)}
        self.opt = {g: torch.optim.Adam([v for k, v in self.p.items() if k.startswith(g + ".")], lr=lr, weight_decay=0.0)
                    for g in OPT_GROUPS}

    def state_dict(self):
        sd = {k: v.detach().clone() for k, v in self.p.items()}
        sd.update({k: v.clone() for k, v in self.tgt.items()})
        return sd

    def _target(self, prefix):
        return {TARGETS[prefix] + k[len(prefix):]: v for k, v in self.tgt.items() if k.startswith(prefix)}

    def act(self, obs, step, eval_mode, num_expl_steps=2000):  # drqv2.py:270-282
        with torch.no_grad():
            state = conv_encoder(self.p, "encoder", torch.as_tensor(obs).unsqueeze(0))
            mu = actor_dist(self.p, state)
            if eval_mode:
                return mu.numpy()[0]
            action = trunc_sample(mu, self.sched(step), None)
            if step < num_expl_steps:
                action.uniform_(-1.0, 1.0)
            return action.numpy()[0]

    def update(self, batch: PixelBatch, step):  # drqv2.py:313-461
        if step % self.up_every != 0:
            return {}
        img, next_img = torch.from_numpy(batch.img), torch.from_numpy(batch.next_img)
        img_step1 = torch.from_numpy(batch.next_img_step)
        action, reward, discount = (torch.from_numpy(x) for x in (batch.action, batch.reward, batch.discount))
        n = img.shape[0]
        img = aug(img.float(), draw_shift(n))
        img_step1 = img_step1[:, -3:, :, :]  # no aug for the predicted frame (uint8: normalised by true division)
        next_img = aug(next_img.float(), draw_shift(n))
        p = self.p
        state = conv_encoder(p, "encoder", img)
        state_step1 = conv_encoder(p, "predict_encoder", img_step1)
        enc_in = torch.cat([state, action, state_step1], dim=-1)
        m, ls = gaussian(p, "feat_encoder", enc_in)
        z = m + ls.exp() * torch.randn(m.shape)  # Normal.rsample
        x = F.relu(F.linear(z, p["feat_decoder.l1.weight"], p["feat_decoder.l1.bias"]))
        x = F.relu(F.linear(x, p["feat_decoder.l2.weight"], p["feat_decoder.l2.bias"]))
        s_hat = F.linear(x, p["feat_decoder.state_linear.weight"], p["feat_decoder.state_linear.bias"])
        r_hat = F.linear(x, p["feat_decoder.reward_linear.weight"], p["feat_decoder.reward_linear.bias"])
        pred = decoder(p, s_hat)
        s_loss = F.l1_loss(pred, img_step1 / 255.0 - 0.5) * 10.0
        r_loss = F.mse_loss(r_hat, reward)
        ml_loss = r_loss + s_loss
        mean1, log_std1 = gaussian(p, "feat_encoder", enc_in)
        sa = torch.cat([state, action], dim=-1)
        mean2, log_std2 = gaussian(p, "feat_f", sa)
        var1, var2 = (2 * log_std1).exp(), (2 * log_std2).exp()
        kl_loss = (log_std2 - log_std1 + 0.5 * (var1 + (mean1 - mean2) ** 2) / var2 - 0.5).mean()
        ae_loss = (ml_loss * self.mse_w + kl_loss) * self.vae_w
        std = self.sched(step)
        with torch.no_grad():
            next_state = conv_encoder(self._target("encoder_target."), "encoder", next_img)
            next_action = trunc_sample(actor_dist(p, next_state), std, self.clip)
            nm, nls = gaussian(self._target("feat_f_target."), "feat_f", torch.cat([next_state, next_action], dim=-1))
            tq1, tq2 = noisy_critic(self._target("critic_target."), "critic", nm, nls, self.c_noise)
            target_q = reward + discount * torch.min(tq1, tq2)
        mean, log_std = gaussian(p, "feat_f", sa)
        q1, q2 = noisy_critic(p, "critic", mean, log_std, self.c_noise)
        critic_loss = F.smooth_l1_loss(q1, target_q) + F.smooth_l1_loss(q2, target_q)
        loss = critic_loss + ae_loss
        groups = ("encoder", "decoder", "predict_encoder", "feat_encoder", "feat_decoder", "feat_f", "critic")
        for g in groups:
            self.opt[g].zero_grad(set_to_none=True)
        loss.backward()
        # gradients of this update, for gradient-level parity tests (the actor step below also back-propagates into
        # feat_f and critic, whose .grad is then no longer the model step's)
        self.last_grads = {k: v.grad.detach().clone() for k, v in p.items() if v.grad is not None and not k.startswith("actor.")}
        for g in groups:
            self.opt[g].step()
        # ---- update_actor(state.detach(), step), drqv2.py:284-311
        obs = state.detach()
        mu = actor_dist(p, obs)
        a = trunc_sample(mu, std, self.clip)
        am, als = gaussian(p, "feat_f", torch.cat([obs, a], dim=-1))
        aq1, aq2 = noisy_critic(p, "critic", am, als, self.c_noise)
        actor_loss = -torch.min(aq1, aq2).mean()
        self.opt["actor"].zero_grad(set_to_none=True)
        actor_loss.backward()
        self.last_grads.update({k: v.grad.detach().clone() for k, v in p.items() if k.startswith("actor.")})
        self.opt["actor"].step()
        with torch.no_grad():  # soft updates, agent_utils.py:42-45
            for k, t in self.tgt.items():
                src = next(TARGETS[pre] + k[len(pre):] for pre in TARGETS if k.startswith(pre))
                t.copy_(self.tau * self.p[src] + (1 - self.tau) * t)
        return {"actor_loss": actor_loss.item(), "critic_loss": critic_loss.item(), "s_loss": s_loss.item(),
                "r_loss": r_loss.item(), "kl_loss": kl_loss.item(), "critic_q1": q1.mean().item(),
                "critic_q2": q2.mean().item(), "critic_target_q": target_q.mean().item()}
