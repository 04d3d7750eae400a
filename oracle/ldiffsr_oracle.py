"""CPU restatement of the latent Diff-SR DrQ-v2 pixel update (reference: agent/diffsrdrq/latent_diff_sr.py:234-390
`LatentDiffSRDrQv2.train_step` with `ae_step`, `score_step`, `critic_step`, `actor_step`, `update_target`; networks
network_arch/vae_1d.py:96-195, score_idql.py:9-70,125-197, latent_diff_sr.py:84-141) -- TEST INFRASTRUCTURE ONLY.

Groundwork for SURVEY.md 8a row a17 (second half): the CUDA path for this agent is not built yet; this file and its
fixture (tests/golden/ldiffsr_*.npz, generated from the real reference class by oracle/make_golden_ldiffsr.py) pin the
arithmetic the next round has to reproduce.

Configuration path restated: configs/latent_diff_sr.yaml -- use_repr_target, back_critic_grad, critic_loss mse,
reg_coef 0, grad_norm null, extra_repr_step 1, do_scale false, bn_dim set.  The online score network runs in training
mode (dropout 0.1 at the head of every MLPResNet block), the targets in eval mode (`make_target` calls .eval()).
RNG order on torch's CPU default generator: shift draw (img_stack), shift draw (next_img_stack), randn[4B, L] (posterior
sample of the 3B + B frames), randint(0, num_noises, [B]), randn[B, L] (diffusion noise), dropout masks of psi (one per
block) and zeta (one per block), dropout masks of psi again (critic step), _standard_normal[B, A] (next action),
_standard_normal[B, A] (actor step).
"""
from __future__ import annotations

import collections
import math

import numpy as np
import torch
import torch.nn.functional as F

from .drq_oracle import PixelBatch, aug, draw_shift, schedule, synthetic_pixel_batch  # noqa: F401  (shared pieces)

REPR = 32 * 35 * 35
Dims = collections.namedtuple("Dims", "A L feat bn psi_h psi_d zeta_h zeta_d H")


def layer_table(d: Dims):
    t = []
    for i, cin in zip((0, 2, 4, 6), (3, 32, 32, 32)):
        t += [(f"vae.encoder.convs.{i}.weight", (32, cin, 3, 3)), (f"vae.encoder.convs.{i}.bias", (32,))]
    t += [("vae.encoder.fc.weight", (d.L, REPR)), ("vae.encoder.fc.bias", (d.L,)), ("vae.encoder.ln.weight", (d.L,)),
          ("vae.encoder.ln.bias", (d.L,)), ("vae.encoder.out.weight", (2 * d.L, d.L)), ("vae.encoder.out.bias", (2 * d.L,)),
          ("vae.decoder.fc.weight", (REPR, d.L)), ("vae.decoder.fc.bias", (REPR,))]
    for i in (0, 2, 4, 6):  # ConvTranspose2d weights are [in, out, kh, kw]
        t += [(f"vae.decoder.deconvs.{i}.weight", (32, 32, 3, 3)), (f"vae.decoder.deconvs.{i}.bias", (32,))]
    t += [("vae.decoder.deconvs.8.weight", (3, 32, 3, 3)), ("vae.decoder.deconvs.8.bias", (3,))]

    def lin(name, o, i):
        return [(name + ".weight", (o, i)), (name + ".bias", (o,))]

    def ln(name, n):
        return [(name + ".weight", (n,)), (name + ".bias", (n,))]

    def resnet(name, blocks, i, o, h):
        r = lin(name + ".fc", h, i)
        for b in range(blocks):
            r += ln(f"{name}.blocks.{b}.layer_norm", h) + lin(f"{name}.blocks.{b}.fc1", 4 * h, h)
            r += lin(f"{name}.blocks.{b}.fc2", h, 4 * h) + lin(f"{name}.blocks.{b}.residual", h, h)
        return r + lin(name + ".out_fc", o, h)
    t += lin("score.psi_bottleneck1.0", d.bn, 3 * d.L) + ln("score.psi_bottleneck1.1", d.bn)
    t += lin("score.psi_bottleneck2.0", d.bn, d.A) + ln("score.psi_bottleneck2.1", d.bn)
    t += resnet("score.psi", d.psi_d, 2 * d.bn, d.feat, d.psi_h)
    t += resnet("score.zeta", d.zeta_d, d.L + d.L // 2, d.L * d.feat, d.zeta_h)
    t += lin("actor.trunk.0", d.bn, 3 * d.L) + ln("actor.trunk.1", d.bn)
    t += lin("actor.policy.0", d.H, d.bn) + lin("actor.policy.2", d.H, d.H) + lin("actor.policy.4", d.A, d.H)
    t += ln("critic.ln", d.feat)
    for a, b, c in (("l1", "l2", "l3"), ("l4", "l5", "l6")):
        t += lin("critic." + a, d.H, d.feat) + lin("critic." + b, d.H, d.H) + lin("critic." + c, 1, d.H)
    return t


def is_layernorm(name):
    return any(s in name for s in (".ln.", ".layer_norm.", "psi_bottleneck1.1.", "psi_bottleneck2.1.", "trunk.1."))


def init_state(d: Dims, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = collections.OrderedDict()
    for name, shape in layer_table(d):
        u = torch.rand(shape, generator=g) * 2 - 1
        if is_layernorm(name):
            sd[name] = 1.0 + 0.1 * u if name.endswith("weight") else 0.1 * u
        else:
            w = shape if name.endswith("weight") else sd[name[:-4] + "weight"].shape
            sd[name] = u / math.sqrt(int(np.prod(w[1:])))
    return sd


def lin(p, n, x):
    return F.linear(x, p[n + ".weight"], p[n + ".bias"])


def lnorm(p, n, x):
    return F.layer_norm(x, (x.shape[-1],), p[n + ".weight"], p[n + ".bias"], 1e-5)


def vae_encode(p, pre, frames):  # vae_1d.Encoder.forward (:118-131) -> (mean, logvar clamped, std, var) (:24-35)
    h = frames / 255.0 - 0.5
    h = F.relu(F.conv2d(h, p[pre + "encoder.convs.0.weight"], p[pre + "encoder.convs.0.bias"], stride=2))
    for i in (2, 4, 6):
        h = F.relu(F.conv2d(h, p[f"{pre}encoder.convs.{i}.weight"], p[f"{pre}encoder.convs.{i}.bias"]))
    h = lnorm(p, pre + "encoder.ln", lin(p, pre + "encoder.fc", h.reshape(h.shape[0], -1)))
    h = lin(p, pre + "encoder.out", h * torch.sigmoid(h))  # swish
    mean, logvar = torch.chunk(h, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    return mean, logvar, torch.exp(0.5 * logvar), torch.exp(logvar)


def vae_decode(p, pre, z):  # vae_1d.Decoder.forward (:156-160)
    h = F.relu(lin(p, pre + "decoder.fc", z)).view(-1, 32, 35, 35)
    for i in (0, 2, 4):
        h = F.relu(F.conv_transpose2d(h, p[f"{pre}decoder.deconvs.{i}.weight"], p[f"{pre}decoder.deconvs.{i}.bias"]))
    h = F.relu(F.conv_transpose2d(h, p[pre + "decoder.deconvs.6.weight"], p[pre + "decoder.deconvs.6.bias"], stride=2,
                                  output_padding=1))
    return F.conv2d(h, p[pre + "decoder.deconvs.8.weight"], p[pre + "decoder.deconvs.8.bias"], padding=1)


def mlp_resnet(p, name, x, blocks, training):  # score_idql.MLPResNet (:45-70); the residual Linear is never reached
    x = lin(p, name + ".fc", x)
    for b in range(blocks):
        r = x
        h = F.dropout(x, 0.1, training)
        h = lnorm(p, f"{name}.blocks.{b}.layer_norm", h)
        h = lin(p, f"{name}.blocks.{b}.fc2", F.mish(lin(p, f"{name}.blocks.{b}.fc1", h)))
        x = r + h
    return lin(p, name + ".out_fc", F.mish(x))


def forward_psi(p, pre, d, state, action, training):  # score_idql.py:172-177
    s = torch.tanh(lnorm(p, pre + "psi_bottleneck1.1", lin(p, pre + "psi_bottleneck1.0", state.reshape(state.shape[0], -1))))
    a = torch.tanh(lnorm(p, pre + "psi_bottleneck2.1", lin(p, pre + "psi_bottleneck2.0", action)))
    return mlp_resnet(p, pre + "psi", torch.cat([s, a], dim=-1), d.psi_d, training)


def time_embed(t, dim):  # score_mlp.SinusoidalPosEmb (:94-106)
    half = dim // 2
    emb = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
    emb = t * emb[None, :]
    return torch.cat((emb.sin(), emb.cos()), dim=-1)


def forward_score(p, pre, d, x_pert, t_idx, psi, training):  # score_idql.py:179-193
    z = mlp_resnet(p, pre + "zeta", torch.cat([x_pert, time_embed(t_idx[..., None], d.L // 2)], dim=-1), d.zeta_d, training)
    score = torch.bmm(psi.unsqueeze(1), z.reshape(-1, d.feat, d.L)).squeeze()
    return score / d.feat


def rff_critic(p, pre, x):  # latent_diff_sr.RFFCritic (:117-141)
    x = lnorm(p, pre + "ln", x)
    q1 = lin(p, pre + "l3", F.elu(lin(p, pre + "l2", torch.sin(lin(p, pre + "l1", x)))))
    q2 = lin(p, pre + "l6", F.elu(lin(p, pre + "l5", torch.sin(lin(p, pre + "l4", x)))))
    return torch.stack([q1, q2], dim=0)


def actor_mu(p, x):
    h = torch.tanh(lnorm(p, "actor.trunk.1", lin(p, "actor.trunk.0", x)))
    return torch.tanh(lin(p, "actor.policy.4", F.relu(lin(p, "actor.policy.2", F.relu(lin(p, "actor.policy.0", h))))))


def trunc_sample(mu, std, clip):  # latent_diff_sr.TruncatedNormal.sample (:72-82)
    eps = torch.normal(torch.zeros(mu.shape), torch.ones(mu.shape)) * std
    if clip is not None:
        eps = torch.clamp(eps, -clip, clip)
    x = mu + eps
    return x - x.detach() + torch.clamp(x, -1.0 + 1e-6, 1.0 - 1e-6).detach()


TARGETS = {"critic_target.": "critic.", "vae_target.": "vae.", "score_target.": "score."}


class OracleLatentDiffSR:
    def __init__(self, dims: Dims, state, *, ae_lr=3e-4, score_lr=3e-4, actor_lr=1e-4, critic_lr=1e-4, tau=0.01, update_every=2,
                 kl_coef=1.0, ae_coef=1.0, repr_coef=1.0, num_noises=1000, noise_param1=1e-4, noise_param2=0.02,
                 stddev_schedule="linear(1.0,0.1,500000)", stddev_clip=0.3):
        self.d, self.tau, self.update_every = dims, tau, update_every
        self.kl_coef, self.ae_coef, self.repr_coef, self.clip = kl_coef, ae_coef, repr_coef, stddev_clip
        self.num_noises, self.sched = num_noises, schedule(stddev_schedule)
        betas = np.linspace(noise_param1, noise_param2, num_noises)  # helper_functions/util.py:118-134
        self.alphabars = torch.as_tensor(np.cumprod(1 - betas, axis=0), dtype=torch.float32)[..., None]
        self.p = {k: v.clone().requires_grad_() for k, v in state.items()}
        self.tgt = {t + k[len(s):]: v.detach().clone() for t, s in TARGETS.items() for k, v in self.p.items()
                    if k.startswith(s)}
        grp = lambda pre: [v for k, v in self.p.items() if k.startswith(pre)]
        self.opt = {"vae": torch.optim.Adam(grp("vae."), lr=ae_lr), "score": torch.optim.AdamW(grp("score."), lr=score_lr),
                    "actor": torch.optim.Adam(grp("actor."), lr=actor_lr), "critic": torch.optim.Adam(grp("critic."), lr=critic_lr)}
        self._step = 1

    def state_dict(self):
        sd = {k: v.detach().clone() for k, v in self.p.items()}
        sd.update({k: v.clone() for k, v in self.tgt.items()})
        return sd

    def train_step(self, batch: PixelBatch, step):  # latent_diff_sr.py:306-353
        self._step += 1
        if self._step % self.update_every != 0:
            return {}
        d, p, t = self.d, self.p, self.tgt
        img, next_img = torch.from_numpy(batch.img), torch.from_numpy(batch.next_img)
        action, reward, discount = (torch.from_numpy(x) for x in (batch.action, batch.reward, batch.discount))
        B = img.shape[0]
        img = aug(img.float(), draw_shift(B)).detach()
        next_img = aug(next_img.float(), draw_shift(B)).detach()
        step1 = torch.from_numpy(batch.next_img_step)[:, -3:].float()
        # ---- ae_step (:234-259): every frame of the stack and the one-step-ahead frame through the per-frame VAE
        frames = torch.cat([img.view(B * 3, 3, 84, 84), step1], dim=0)
        mean, logvar, std, var = vae_encode(p, "vae.", frames)
        z_all = mean + std * torch.randn(mean.shape)
        pred = vae_decode(p, "vae.", z_all)
        recon_loss = F.mse_loss(pred, frames / 255.0 - 0.5, reduction="sum") / pred.shape[0]
        kl_loss = (0.5 * (mean.pow(2) + var - 1.0 - logvar).sum(-1)).mean()
        ae_loss = recon_loss + self.kl_coef * kl_loss
        latent, next_latent_step = torch.split(z_all, [B * 3, B], dim=0)
        latent = latent.reshape(B, -1)
        latent_mode = torch.split(mean, [B * 3, B], dim=0)[0].reshape(B, -1)
        # ---- score_step (:275-304)
        noise_idx = torch.randint(0, self.num_noises, (B,))
        ab = self.alphabars[noise_idx]
        noise = torch.randn_like(next_latent_step)
        pert = ab.sqrt() * next_latent_step + (1 - ab).sqrt() * noise
        psi = forward_psi(p, "score.", d, latent, action, True)
        score = forward_score(p, "score.", d, pert, noise_idx, psi, True)
        score_loss = (score * (1 - ab).sqrt() + noise).pow(2).sum(1).mean()
        # ---- critic_step (:355-379), back_critic_grad
        feature = forward_psi(p, "score.", d, latent_mode, action, True)
        std_now = self.sched(step)
        with torch.no_grad():
            nm = vae_encode(t, "vae_target.", next_img.view(B * 3, 3, 84, 84))[0].reshape(B, -1)
            next_action = trunc_sample(actor_mu(p, nm), std_now, self.clip)
            next_feature = forward_psi(t, "score_target.", d, nm, next_action, False)
            q_target = reward + discount * rff_critic(t, "critic_target.", next_feature).min(0)[0]
        q_pred = rff_critic(p, "critic.", feature)
        critic_loss = F.mse_loss(q_pred, q_target.unsqueeze(0).repeat(2, 1, 1))
        loss = (ae_loss * self.ae_coef + score_loss) * self.repr_coef + critic_loss
        for g in ("vae", "score", "critic"):
            self.opt[g].zero_grad(set_to_none=True)
        loss.backward()
        self.last_grads = {k: v.grad.detach().clone() for k, v in p.items() if v.grad is not None}
        for g in ("vae", "score", "critic"):
            self.opt[g].step()
        # ---- actor_step (:381-390) on the detached posterior mode, through the frozen target psi and the new critic
        lat = latent_mode.detach()
        a = trunc_sample(actor_mu(p, lat), std_now, self.clip)
        actor_loss = -rff_critic(p, "critic.", forward_psi(t, "score_target.", d, lat, a, False)).min(0)[0].mean()
        self.opt["actor"].zero_grad(set_to_none=True)
        actor_loss.backward()
        self.opt["actor"].step()
        with torch.no_grad():  # update_target (:135-139), helper_functions/util.py:136-138
            for k, tv in t.items():
                src = next(TARGETS[pre] + k[len(pre):] for pre in TARGETS if k.startswith(pre))
                tv.copy_(tv * (1.0 - self.tau) + p[src] * self.tau)
        return {"loss/recon_loss": recon_loss.item(), "loss/kl_loss": kl_loss.item(), "loss/score_loss": score_loss.item(),
                "loss/reg_loss": 0.0, "info/psi_l1_norm": psi.abs().mean().item(), "loss/actor_loss": actor_loss.item(),
                "loss/critic_loss": critic_loss.item(), "info/q_pred": q_pred.mean().item(),
                "info/q_target": q_target.mean().item(), "info/reward": reward.mean().item(),
                "info/latent_mean": latent.mean().item(), "info/latent_std": latent.std().item(),
                "info/latent_l1_norm": latent.abs().mean().item(), "info/latent_dist_mean": mean.mean().item(),
                "info/latent_dist_std": std.mean().item(), "info/policy_std": float(std_now),
                "info/vae_grad_norm": 0.0, "info/score_grad_norm": 0.0}
