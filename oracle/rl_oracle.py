"""CPU restatement of the rl-rep `agent.train()` update step (reference: haotiansun14/rl-rep).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Plain PyTorch fp32 on the CPU (the reference's own
arithmetic lives in un-vendored PyTorch: F.linear, F.elu, autograd, torch.optim.Adam -- SURVEY.md 8c), written
functionally over a flat {state_dict name -> tensor} table instead of nn.Module classes.  Every function cites
the reference file:line it follows (paths relative to the reference root).

Pinning: the reference ships no tests or golden vectors ("parity unpinned" by the reference itself).  This
restatement is pinned operationally instead: oracle/make_golden.py imports the real reference classes in the
build container, runs them on seeded inputs, and commits the outputs under tests/golden/;
tests/test_oracle_golden.py checks this file against those fixtures.

Random numbers are consumed from the same global generators, in the same order, as the reference does
(legacy numpy RNG for replay indices, torch CPU default generator for every epsilon -- SURVEY.md A.5), so
seeding both identically yields identical draws.
"""
from __future__ import annotations

import collections
import copy
import math

import numpy as np
import torch
import torch.nn.functional as F

from bench_data import Batch  # noqa: E402


# Replay ring, synthetic transitions, layer tables and deterministic initial weights live in bench_data.py (data only;
# bench.py's GPU arm uses them without importing this module).
from bench_data import HostRing, init_state, layer_table, synthetic_ring  # noqa: E402,F401

# --------------------------------------------------------------------------------------------------
# Shared SAC maths
# --------------------------------------------------------------------------------------------------
def lin(p, name, x):
    return F.linear(x, p[name + ".weight"], p[name + ".bias"])


def actor_head(p, obs):
    """DiagGaussianActor.forward (agent/sac/actor.py:76-91): trunk -> (mu, log_std) with
    log_std = -5 + 3.5 (tanh(raw) + 1); returns (mu, std)."""
    h = F.elu(lin(p, "actor.trunk.0", obs))
    h = F.elu(lin(p, "actor.trunk.2", h))
    mu, raw = lin(p, "actor.trunk.4", h).chunk(2, dim=-1)
    log_std = torch.tanh(raw)
    log_std = -5.0 + 0.5 * (2.0 - (-5.0)) * (log_std + 1)
    return mu, log_std.exp()


def squash_sample(mu, std, eps):
    """SquashedNormal rsample + log_prob (actor.py:16-60 over torch.distributions): u = mu + eps*std,
    a = tanh(u), log pi = sum_A [ N(u; mu, std).log_prob - 2 (log 2 - u - softplus(-2u)) ]."""
    u = mu + eps * std
    a = torch.tanh(u)
    base = -((u - mu) ** 2) / (2 * std ** 2) - std.log() - math.log(math.sqrt(2 * math.pi))
    ladj = 2.0 * (math.log(2.0) - u - F.softplus(-2.0 * u))
    return a, (-ladj + base).sum(-1, keepdim=True)


class OracleSAC:
    """agent/sac/sac_agent.py:16-188.  Subclasses override the representation part."""
    alg = "sac"

    def __init__(self, state_dim, action_dim, state: dict, *, lr=3e-4, discount=0.99, target_update_period=2,
                 tau=0.005, alpha=0.1, auto_entropy_tuning=True, hidden_dim=1024, action_range=(-1.0, 1.0), **cfg):
        self.S, self.A = state_dim, action_dim
        self.cfg = dict(cfg, hidden_dim=hidden_dim)
        self.discount, self.tau, self.period = discount, tau, target_update_period
        self.learn_alpha = auto_entropy_tuning
        self.action_range = action_range
        self.steps = 0
        self.p = {k: v.detach().clone().float().requires_grad_(True) for k, v in state.items()}
        self.log_alpha = torch.tensor(np.log(alpha), requires_grad=True)  # float64, sac_agent.py:66
        self.target_entropy = -action_dim
        self.critic_target = self._clone("critic.")  # sac_agent.py:56
        self._make_optimizers(lr)

    # -- plumbing
    def _group(self, prefix):
        return [v for k, v in self.p.items() if k.startswith(prefix)]

    def _clone(self, prefix):
        return {k: v.detach().clone() for k, v in self.p.items() if k.startswith(prefix)}

    def _make_optimizers(self, lr):  # sac_agent.py:71-81
        self.actor_opt = torch.optim.Adam(self._group("actor."), lr=lr, betas=[0.9, 0.999])
        self.critic_opt = torch.optim.Adam(self._group("critic."), lr=lr, betas=[0.9, 0.999])
        self.alpha_opt = torch.optim.Adam([self.log_alpha], lr=lr, betas=[0.9, 0.999])

    @property
    def alpha(self):
        return self.log_alpha.exp()

    @staticmethod
    def _polyak(src: dict, dst: dict, tau):  # sac_agent.py:99-102
        with torch.no_grad():
            for k in dst:
                dst[k].copy_(tau * src[k].data + (1 - tau) * dst[k])

    def update_target(self):
        if self.steps % self.period == 0:
            self._polyak(self.p, self.critic_target, self.tau)

    def _note_grads(self, opt):
        """Test aid: keeps the gradient each parameter had at its most recent optimiser step (`last_step_grads`), so parity
        tests can tell which elements' Adam steps are ill-conditioned (|g| ~ 0)."""
        if not hasattr(self, "last_step_grads"):
            self.last_step_grads, self._names_by_id = {}, {}
        if len(self._names_by_id) != len(self.p):
            self._names_by_id = {id(v): k for k, v in self.p.items()}
        for group in opt.param_groups:
            for q in group["params"]:
                name = self._names_by_id.get(id(q))
                if name is not None and q.grad is not None:
                    self.last_step_grads[name] = q.grad.detach().clone()

    def state_dict(self):
        sd = {k: v.detach().clone() for k, v in self.p.items()}
        sd.update({"critic_target." + k[len("critic."):]: v.clone() for k, v in self.critic_target.items()})
        sd["log_alpha"] = self.log_alpha.detach().clone()
        return sd

    # -- inference (sac_agent.py:89-96)
    def select_action(self, state, explore=False):
        with torch.no_grad():
            obs = torch.FloatTensor(state).unsqueeze(0)
            mu, std = actor_head(self.p, obs)
            a = torch.tanh(mu + torch.randn(mu.shape) * std) if explore else torch.tanh(mu)
            return a.clamp(*self.action_range)[0].numpy()

    # -- critic on raw (s, a): DoubleQCritic.forward (agent/sac/critic.py:26-36)
    @staticmethod
    def _q_mlp(p, prefix, x):
        h = F.elu(lin(p, prefix + ".0", x))
        h = F.elu(lin(p, prefix + ".2", h))
        return lin(p, prefix + ".4", h)

    def _twin_q(self, p, obs, act):
        x = torch.cat([obs, act], dim=-1)
        return self._q_mlp(p, "critic.Q1", x), self._q_mlp(p, "critic.Q2", x)

    def critic_step(self, b: Batch):  # sac_agent.py:105-135
        mu, std = actor_head(self.p, b.next_state)
        a2, logp = squash_sample(mu, std, torch.randn(mu.shape))
        tq1, tq2 = self._twin_q(self.critic_target, b.next_state, a2)
        target_v = torch.min(tq1, tq2) - self.alpha.detach() * logp
        target_q = (b.reward + (1.0 - b.done) * self.discount * target_v).detach()
        q1, q2 = self._twin_q(self.p, b.state, b.action)
        loss = F.mse_loss(q1, target_q) + F.mse_loss(q2, target_q)
        self.critic_opt.zero_grad()
        loss.backward()
        self._note_grads(self.critic_opt)
        self.critic_opt.step()
        return {"q_loss": loss.item(), "q1": q1.mean().item(), "q2": q1.mean().item()}  # q2 := mean(Q1), :134

    def _actor_q(self, obs, act):
        return self._twin_q(self.p, obs, act)

    _alpha_detached_in_actor_loss = True  # sac_agent.py:147; ctrlsac/vlsac/spedersac use the live alpha

    def update_actor_and_alpha(self, b: Batch):  # sac_agent.py:138-166
        mu, std = actor_head(self.p, b.state)
        a, logp = squash_sample(mu, std, torch.randn(mu.shape))
        q1, q2 = self._actor_q(b.state, a)
        alpha = self.alpha.detach() if self._alpha_detached_in_actor_loss else self.alpha
        actor_loss = (alpha * logp - torch.min(q1, q2)).mean()
        self.actor_opt.zero_grad()
        actor_loss.backward()
        self._note_grads(self.actor_opt)
        self.actor_opt.step()
        info = {"actor_loss": actor_loss.item()}
        if self.learn_alpha:
            self.alpha_opt.zero_grad()
            alpha_loss = (self.alpha * (-logp - self.target_entropy).detach()).mean()
            alpha_loss.backward()
            self._note_grads(self.alpha_opt)
            self.alpha_opt.step()
            info["alpha_loss"] = alpha_loss.item()
            info["alpha"] = self.alpha.item()
        return info

    def train(self, buffer, batch_size):  # sac_agent.py:169-188
        self.steps += 1
        b = buffer.sample(batch_size)
        info = self.critic_step(b)
        info.update(self.update_actor_and_alpha(b))
        self.update_target()
        return info


# --------------------------------------------------------------------------------------------------
# CTRL-SAC (agent/ctrlsac/ctrlsac_agent.py)
# --------------------------------------------------------------------------------------------------
def ctrl_phi(p, s, a, pre="phi"):  # Phi.forward, ctrlsac_agent.py:72-77
    z = F.elu(lin(p, pre + ".l1", torch.cat([s, a], dim=-1)))
    z = F.elu(lin(p, pre + ".l2", z))
    return lin(p, pre + ".l3", z)


def ctrl_mu(p, s2):  # Mu.forward, ctrlsac_agent.py:96-102 (bounded by tanh)
    z = F.elu(lin(p, "mu.l1", s2))
    z = F.elu(lin(p, "mu.l2", z))
    return torch.tanh(lin(p, "mu.l3", z))


def ctrl_critic(p, z):  # Critic.forward, ctrlsac_agent.py:40-52: one ELU hidden layer per head
    return lin(p, "critic.l2", F.elu(lin(p, "critic.l1", z))), lin(p, "critic.l5", F.elu(lin(p, "critic.l4", z)))


class OracleCTRLSAC(OracleSAC):
    alg = "ctrlsac"
    _alpha_detached_in_actor_loss = False  # ctrlsac_agent.py:308

    def __init__(self, state_dim, action_dim, state, *, lr=1e-4, hidden_dim=1024, feature_tau=0.005,
                 feature_dim=2048, use_feature_target=True, extra_feature_steps=1, as_written=True, **kw):
        self.feature_tau, self.use_ft, self.K = feature_tau, use_feature_target, extra_feature_steps + 1
        # as_written=True builds the logits with the reference's [B,1,D]*[1,B,D] broadcast (:229);
        # False restates that one line as a matmul (needed for B >= 2048).
        self.as_written = as_written
        super().__init__(state_dim, action_dim, state, lr=lr, hidden_dim=hidden_dim, feature_dim=feature_dim, **kw)
        self.phi_target = self._clone("phi.")  # :170-171

    def _make_optimizers(self, lr):  # ctrlsac_agent.py:176-203
        feat = self._group("phi.") + self._group("mu.") + self._group("theta.")
        self.feature_opt = torch.optim.Adam(feat, weight_decay=0, lr=lr)
        self.actor_opt = torch.optim.Adam(self._group("actor."), weight_decay=0, lr=lr / 3, betas=[0.9, 0.999])
        self.alpha_opt = torch.optim.Adam([self.log_alpha], lr=lr / 3, betas=[0.9, 0.999])
        self.critic_opt = torch.optim.Adam(self._group("critic."), weight_decay=0, lr=lr, betas=[0.9, 0.999])

    def state_dict(self):
        sd = super().state_dict()
        sd.update({"phi_target." + k[len("phi."):]: v.clone() for k, v in self.phi_target.items()})
        return sd

    def feature_step(self, b: Batch):  # ctrlsac_agent.py:213-251
        z_phi = ctrl_phi(self.p, b.state, b.action)
        z_mu = ctrl_mu(self.p, b.next_state)
        if self.as_written:
            logits = (z_phi[:, None, :] * z_mu[None, :, :]).sum(-1)
        else:
            logits = z_phi @ z_mu.t()
        labels = torch.eye(b.state.shape[0])
        model_loss = F.cross_entropy(logits, labels)  # soft-label CE == -mean(diag(log_softmax))
        r_loss = 0.5 * F.mse_loss(lin(self.p, "theta.l", z_phi), b.reward).mean()
        loss = model_loss + r_loss
        self.feature_opt.zero_grad()
        loss.backward()
        self._note_grads(self.feature_opt)
        self.feature_opt.step()
        return {"total_loss": loss.item(), "model_loss": model_loss.item(), "r_loss": r_loss.item()}

    def critic_step(self, b: Batch):  # ctrlsac_agent.py:257-293
        # frozen_phi_target is loaded from *phi* (:346), i.e. the live phi after the feature loop.
        with torch.no_grad():
            mu, std = actor_head(self.p, b.next_state)
            a2, logp2 = squash_sample(mu, std, torch.randn(mu.shape))
            z = ctrl_phi(self.p, b.state, b.action)
            z2 = ctrl_phi(self.p, b.next_state, a2)
            nq1, nq2 = ctrl_critic(self.critic_target, z2)
            next_q = torch.min(nq1, nq2) - self.alpha * logp2
            target_q = b.reward + (1.0 - b.done) * self.discount * next_q
        q1, q2 = ctrl_critic(self.p, z)
        q1_loss, q2_loss = F.mse_loss(target_q, q1), F.mse_loss(target_q, q2)
        self.critic_opt.zero_grad()
        (q1_loss + q2_loss).backward()
        self._note_grads(self.critic_opt)
        self.critic_opt.step()
        return {"q1_loss": q1_loss.item(), "q2_loss": q2_loss.item(), "q1": q1.mean().item(), "q2": q2.mean().item()}

    def _actor_q(self, obs, act):  # ctrlsac_agent.py:303-305; frozen_phi == phi after :344
        frozen = {k: v.detach() for k, v in self.p.items() if k.startswith("phi.")}
        return ctrl_critic(self.p, ctrl_phi(frozen, obs, act))

    def train(self, buffer, batch_size):  # ctrlsac_agent.py:327-362
        self.steps += 1
        for _ in range(self.K):
            b = buffer.sample(batch_size)
            info = self.feature_step(b)
            if self.use_ft:
                self._polyak(self.p, self.phi_target, self.feature_tau)  # :253-255; never read afterwards
        info.update(self.critic_step(b))  # last feature batch
        info.update(self.update_actor_and_alpha(b))
        self.update_target()
        return info


# --------------------------------------------------------------------------------------------------
# VL-SAC / LV-Rep (agent/vlsac/vlsac_agent.py + networks/vae.py)
# --------------------------------------------------------------------------------------------------
LOG_SIG_MAX, LOG_SIG_MIN = 2, -20  # networks/vae.py:9-10


def vae_encoder(p, s, a, s2):  # Encoder.forward, vae.py:36-48
    z = F.relu(lin(p, "encoder.l1", torch.cat([s, a, s2], dim=-1)))
    z = F.relu(lin(p, "encoder.l2", z))
    return lin(p, "encoder.mean_linear", z), torch.clamp(lin(p, "encoder.log_std_linear", z), LOG_SIG_MIN, LOG_SIG_MAX)


def vae_decoder(p, z):  # Decoder.forward, vae.py:79-86
    x = F.relu(lin(p, "decoder.l1", z))
    return lin(p, "decoder.state_linear", x), lin(p, "decoder.reward_linear", x)


def vae_prior(p, s, a, pre="f"):  # GaussianFeature.forward, vae.py:110-120
    z = F.relu(lin(p, pre + ".l1", torch.cat([s, a], dim=-1)))
    z = F.relu(lin(p, pre + ".l2", z))
    return lin(p, pre + ".mean_linear", z), torch.clamp(lin(p, pre + ".log_std_linear", z), LOG_SIG_MIN, LOG_SIG_MAX)


def vl_critic(p, noise, mean, log_std):  # vlsac Critic.forward, vlsac_agent.py:44-63 -- Q2's head is l3 (:61)
    std = log_std.exp()
    B, d = mean.shape
    x = (mean[:, None, :] + std[:, None, :] * noise).reshape(-1, d)
    q1 = F.elu(lin(p, "critic.l1", x)).reshape(B, noise.shape[0], -1).mean(dim=1)
    q1 = lin(p, "critic.l3", F.elu(lin(p, "critic.l2", q1)))
    q2 = F.elu(lin(p, "critic.l4", x)).reshape(B, noise.shape[0], -1).mean(dim=1)
    q2 = lin(p, "critic.l3", F.elu(lin(p, "critic.l5", q2)))
    return q1, q2


class OracleVLSAC(OracleSAC):
    alg = "vlsac"
    _alpha_detached_in_actor_loss = False  # vlsac_agent.py:180

    def __init__(self, state_dim, action_dim, state, *, critic_noise, lr=1e-4, hidden_dim=256, feature_tau=0.001,
                 feature_dim=256, use_feature_target=True, extra_feature_steps=1, **kw):
        self.feature_tau, self.use_ft, self.K = feature_tau, use_feature_target, extra_feature_steps + 1
        self.noise = critic_noise.clone()  # [20, D], drawn once at construction (vlsac_agent.py:30-31)
        super().__init__(state_dim, action_dim, state, lr=lr, hidden_dim=hidden_dim, feature_dim=feature_dim, **kw)
        self.f_target = self._clone("f.")

    def _make_optimizers(self, lr):  # SACAgent's actor/alpha optimisers are kept (lr), vlsac_agent.py:116-123
        super()._make_optimizers(lr)
        feat = self._group("encoder.") + self._group("decoder.") + self._group("f.")
        self.feature_opt = torch.optim.Adam(feat, lr=lr)
        self.critic_opt = torch.optim.Adam(self._group("critic."), lr=lr, betas=[0.9, 0.999])

    def state_dict(self):
        sd = super().state_dict()
        sd.update({"f_target." + k[len("f."):]: v.clone() for k, v in self.f_target.items()})
        return sd

    def feature_step(self, b: Batch):  # vlsac_agent.py:126-162
        mean, log_std = vae_encoder(self.p, b.state, b.action, b.next_state)
        z = mean + torch.randn(mean.shape) * log_std.exp()  # Encoder.sample, vae.py:50-58
        x, r = vae_decoder(self.p, z)
        s_loss = 0.5 * F.mse_loss(x, b.next_state)
        r_loss = 0.5 * F.mse_loss(r, b.reward)
        ml_loss = r_loss + s_loss
        mean1, log_std1 = vae_encoder(self.p, b.state, b.action, b.next_state)  # second forward, :143
        mean2, log_std2 = vae_prior(self.p, b.state, b.action)
        var1, var2 = (2 * log_std1).exp(), (2 * log_std2).exp()
        kl = log_std2 - log_std1 + 0.5 * (var1 + (mean1 - mean2) ** 2) / var2 - 0.5
        loss = (ml_loss + kl).mean()
        self.feature_opt.zero_grad()
        loss.backward()
        self._note_grads(self.feature_opt)
        self.feature_opt.step()
        return {"vae_loss": loss.item(), "ml_loss": ml_loss.mean().item(), "kl_loss": kl.mean().item(),
                "s_loss": s_loss.mean().item(), "r_loss": r_loss.mean().item()}

    def _feat(self):
        if self.use_ft:
            return {"f." + k[len("f."):]: v for k, v in self.f_target.items()}
        return self.p

    def critic_step(self, b: Batch):  # vlsac_agent.py:201-237
        with torch.no_grad():
            mu, std = actor_head(self.p, b.next_state)
            a2, logp2 = squash_sample(mu, std, torch.randn(mu.shape))
            mean, log_std = vae_prior(self._feat(), b.state, b.action)
            nmean, nlog_std = vae_prior(self._feat(), b.next_state, a2)
            nq1, nq2 = vl_critic(self.critic_target, self.noise, nmean, nlog_std)
            next_q = torch.min(nq1, nq2) - self.alpha * logp2
            target_q = b.reward + (1.0 - b.done) * self.discount * next_q
        q1, q2 = vl_critic(self.p, self.noise, mean, log_std)
        q1_loss, q2_loss = F.mse_loss(target_q, q1), F.mse_loss(target_q, q2)
        self.critic_opt.zero_grad()
        (q1_loss + q2_loss).backward()
        self._note_grads(self.critic_opt)
        self.critic_opt.step()
        return {"q1_loss": q1_loss.item(), "q2_loss": q2_loss.item(), "q1": q1.mean().item(), "q2": q2.mean().item()}

    def _actor_q(self, obs, act):  # vlsac_agent.py:173-178
        ft = {k: v.detach() for k, v in self._feat().items() if k.startswith("f.")}
        mean, log_std = vae_prior(ft, obs, act)
        return vl_critic(self.p, self.noise, mean, log_std)

    def train(self, buffer, batch_size):  # vlsac_agent.py:245-273
        self.steps += 1
        for _ in range(self.K):
            b = buffer.sample(batch_size)
            info = self.feature_step(b)
            if self.use_ft:
                self._polyak(self.p, self.f_target, self.feature_tau)
        info.update(self.critic_step(b))
        info.update(self.update_actor_and_alpha(b))
        self.update_target()
        return info


# --------------------------------------------------------------------------------------------------
# SPEDER-SAC (agent/spedersac/spedersac_agent.py)
# --------------------------------------------------------------------------------------------------
def _trunk(p, prefix, x):
    """nn.Sequential built by util.mlp / spedersac mlp (spedersac_agent.py:67-76): Linear, ELU, ..., Linear."""
    idx = sorted({int(k[len(prefix) + 1:].split(".")[0]) for k in p if k.startswith(prefix + ".")})
    for n, i in enumerate(idx):
        x = lin(p, f"{prefix}.{i}", x)
        if n + 1 < len(idx):
            x = F.elu(x)
    return x


def rff_critic(p, z):  # RFFCritic.forward, spedersac_agent.py:38-50: sin -> ELU -> linear, two heads
    q1 = lin(p, "critic.l3", F.elu(lin(p, "critic.l2", torch.sin(lin(p, "critic.l1", z)))))
    q2 = lin(p, "critic.l6", F.elu(lin(p, "critic.l5", torch.sin(lin(p, "critic.l4", z)))))
    return q1, q2


class OracleSPEDERSAC(OracleSAC):
    alg = "spedersac"
    _alpha_detached_in_actor_loss = False  # spedersac_agent.py:272

    def __init__(self, state_dim, action_dim, state, *, phi_and_mu_lr, phi_hidden_dim, phi_hidden_depth, mu_hidden_dim,
                 mu_hidden_depth, critic_and_actor_lr, critic_and_actor_hidden_dim, hidden_dim=1024, feature_tau=0.005,
                 feature_dim=2048, use_feature_target=True, extra_feature_steps=1, **kw):
        self.feature_tau, self.use_ft, self.K = feature_tau, use_feature_target, extra_feature_steps + 1
        self._lrs = (phi_and_mu_lr, critic_and_actor_lr)
        super().__init__(state_dim, action_dim, state, hidden_dim=hidden_dim, feature_dim=feature_dim,
                         phi_hidden_dim=phi_hidden_dim, phi_hidden_depth=phi_hidden_depth, mu_hidden_dim=mu_hidden_dim,
                         mu_hidden_depth=mu_hidden_depth, critic_and_actor_hidden_dim=critic_and_actor_hidden_dim, **kw)
        self.phi_target = self._clone("phi.")

    def _make_optimizers(self, lr):  # spedersac_agent.py:160-179
        flr, clr = self._lrs
        feat = self._group("phi.") + self._group("mu.") + self._group("theta.")
        self.feature_opt = torch.optim.Adam(feat, weight_decay=0, lr=flr)
        self.actor_opt = torch.optim.Adam(self._group("actor."), weight_decay=0, lr=clr, betas=[0.9, 0.999])
        self.alpha_opt = torch.optim.Adam([self.log_alpha], lr=clr, betas=[0.9, 0.999])
        self.critic_opt = torch.optim.Adam(self._group("critic."), weight_decay=0, lr=clr, betas=[0.9, 0.999])

    def state_dict(self):
        sd = super().state_dict()
        sd.update({"phi_target." + k[len("phi."):]: v.clone() for k, v in self.phi_target.items()})
        return sd

    def feature_step(self, b: Batch, b_rand: Batch):  # spedersac_agent.py:181-219
        z_phi = _trunk(self.p, "phi.trunk", torch.cat([b.state, b.action], -1))
        z_phi_r = _trunk(self.p, "phi.trunk", torch.cat([b_rand.state, b_rand.action], -1))
        z_mu = _trunk(self.p, "mu.trunk", b.next_state)
        z_mu_r = _trunk(self.p, "mu.trunk", b_rand.next_state)
        pt1 = -2 * torch.diag(z_phi @ z_mu.T)
        pa = z_phi_r @ z_mu_r.T
        pt2 = pa @ pa.T
        model_loss = 1.0 / pt1.numel() * pt1.sum() + 1.0 / pt2.numel() * pt2.sum()
        r_loss = 0.5 * F.mse_loss(lin(self.p, "theta.l", z_phi), b.reward).mean()
        loss = model_loss + r_loss
        self.feature_opt.zero_grad()
        loss.backward()
        self._note_grads(self.feature_opt)
        self.feature_opt.step()
        return {"total_loss": loss.item(), "model_loss": model_loss.item(), "r_loss": r_loss.item()}

    def critic_step(self, b: Batch):  # spedersac_agent.py:225-257 (live phi under no_grad)
        with torch.no_grad():
            mu, std = actor_head(self.p, b.next_state)
            a2, logp2 = squash_sample(mu, std, torch.randn(mu.shape))
            z = _trunk(self.p, "phi.trunk", torch.cat([b.state, b.action], -1))
            z2 = _trunk(self.p, "phi.trunk", torch.cat([b.next_state, a2], -1))
            nq1, nq2 = rff_critic(self.critic_target, z2)
            next_q = torch.min(nq1, nq2) - self.alpha * logp2
            target_q = b.reward + (1.0 - b.done) * self.discount * next_q
        q1, q2 = rff_critic(self.p, z)
        q1_loss, q2_loss = F.mse_loss(target_q, q1), F.mse_loss(target_q, q2)
        self.critic_opt.zero_grad()
        (q1_loss + q2_loss).backward()
        self._note_grads(self.critic_opt)
        self.critic_opt.step()
        return {"q1_loss": q1_loss.item(), "q2_loss": q2_loss.item(), "q1": q1.mean().item(), "q2": q2.mean().item()}

    def _actor_q(self, obs, act):  # spedersac_agent.py:267-270 -- phi is live here (its grads are discarded)
        return rff_critic(self.p, _trunk(self.p, "phi.trunk", torch.cat([obs, act], -1)))

    def train(self, buffer, batch_size):  # spedersac_agent.py:291-322
        self.steps += 1
        for _ in range(self.K):
            b1 = buffer.sample(batch_size)
            b2 = buffer.sample(batch_size)
            info = self.feature_step(b1, b2)
            if self.use_ft:
                self._polyak(self.p, self.phi_target, self.feature_tau)
        info.update(self.critic_step(b1))
        info.update(self.update_actor_and_alpha(b1))
        self.update_target()
        return info


# --------------------------------------------------------------------------------------------------
# Diff-SR-SAC (agent/diffsrsac/diffsrsac_agent.py)
# --------------------------------------------------------------------------------------------------
def diffsr_alphabars(a=0.3, b=0.1, num=1000):
    """generate_alphabars_and_alphas (diffsrsac_agent.py:178-203): alpha-bar table from the Beta(a,b) CDF,
    clipped to [raw[-2], raw[1]]."""
    from scipy.stats import beta
    raw = 1.0 - beta.cdf(np.linspace(0, 1, num), a, b)
    return torch.tensor(np.clip(raw, a_min=raw[-2], a_max=raw[1])).float()


def diffsr_critic(p, z):  # RFFCritic.forward, diffsrsac_agent.py:60-90 (lambda = 0: the reg term is exactly 0)
    return rff_critic(p, z)


class OracleDIFFSRSAC(OracleSAC):
    alg = "diffsrsac"
    _alpha_detached_in_actor_loss = True  # diffsrsac_agent.py:250

    def __init__(self, state_dim, action_dim, state, *, feature_dim=256, phi_and_nabla_mu_lr=0.003, phi_hidden_dim=256,
                 phi_hidden_depth=1, nabla_mu_hidden_dim=512, nabla_mu_hidden_depth=1, critic_and_actor_lr=3e-4,
                 hidden_dim=256, extra_feature_steps=3, num_noises=1000, DARL_noise_a=0.3, DARL_noise_b=0.1,
                 sigma_scale_factor=0.449, **kw):
        self.K, self.num_noises, self.sigma = extra_feature_steps + 1, num_noises, sigma_scale_factor
        self.alphabars = diffsr_alphabars(DARL_noise_a, DARL_noise_b, num_noises)
        self._flr = phi_and_nabla_mu_lr
        self.D = feature_dim
        super().__init__(state_dim, action_dim, state, lr=critic_and_actor_lr, hidden_dim=hidden_dim,
                         feature_dim=feature_dim, phi_hidden_dim=phi_hidden_dim, phi_hidden_depth=phi_hidden_depth,
                         nabla_mu_hidden_dim=nabla_mu_hidden_dim, nabla_mu_hidden_depth=nabla_mu_hidden_depth, **kw)

    def _make_optimizers(self, lr):
        # SACAgent.__init__ binds critic_optimizer to the *base* DoubleQCritic, which DIFFSRSACAgent then replaces
        # (diffsrsac_agent.py:156-168): the RFF critic is never optimised.  No critic optimiser.
        self.actor_opt = torch.optim.Adam(self._group("actor."), lr=lr, betas=[0.9, 0.999])
        self.alpha_opt = torch.optim.Adam([self.log_alpha], lr=lr, betas=[0.9, 0.999])
        self.phi_opt = torch.optim.Adam(self._group("critic_feed_feature."), lr=self._flr, betas=[0.9, 0.999])
        self.nabla_opt = torch.optim.Adam(self._group("nablamu_net."), lr=self._flr, betas=[0.9, 0.999])

    def _phi(self, p, s, a):
        return _trunk(p, "critic_feed_feature.z_vector", torch.cat([s, a], -1))

    def feature_step(self, b: Batch):  # critic_feeder_feature_step, diffsrsac_agent.py:271-318
        B = b.action.shape[0]
        idx = torch.randint(0, self.num_noises, (B,))
        ab = torch.index_select(self.alphabars, 0, idx).float().reshape(B, 1)
        noise = torch.normal(mean=torch.zeros_like(b.next_state), std=torch.ones_like(b.next_state) * self.sigma)
        s2_pert = torch.sqrt(ab) * b.next_state + torch.sqrt(1.0 - ab) * noise
        target = -(s2_pert - torch.sqrt(ab) * b.next_state)
        phi = self._phi(self.p, b.state, b.action)
        flat = _trunk(self.p, "nablamu_net.Mu_z_by_s_layer", torch.cat([s2_pert, ab], -1))
        score = torch.bmm(phi.unsqueeze(1), flat.reshape(B, self.D, self.S)).squeeze()
        diff = target - (1.0 - ab) * self.sigma * score
        score_loss = ((1 / B) * torch.sum(diff ** 2, dim=list(range(1, diff.dim())))).sum()
        self.nabla_opt.zero_grad()
        self.phi_opt.zero_grad()
        score_loss.backward()
        self._note_grads(self.nabla_opt)
        self.nabla_opt.step()
        self._note_grads(self.phi_opt)
        self.phi_opt.step()
        return {"score_loss": score_loss.item()}

    def critic_step(self, b: Batch):  # diffsrsac_agent.py:205-239 -- losses are computed, nothing is stepped
        with torch.no_grad():
            mu, std = actor_head(self.p, b.next_state)
            a2, logp = squash_sample(mu, std, torch.randn(mu.shape))
            tq1, tq2 = diffsr_critic(self.critic_target, self._phi(self.p, b.next_state, a2))
            target_q = b.reward + (1.0 - b.done) * self.discount * (torch.min(tq1, tq2) - self.alpha.detach() * logp)
            q1, q2 = diffsr_critic(self.p, self._phi(self.p, b.state, b.action))
            noreg = F.mse_loss(q1, target_q) + F.mse_loss(q2, target_q)
        return {"q_loss_reg": noreg.item(), "q_loss_noreg": noreg.item(), "q1": q1.mean().item(),
                "q2": q1.mean().item()}

    def _actor_q(self, obs, act):  # diffsrsac_agent.py:247
        return diffsr_critic(self.p, self._phi(self.p, obs, act))

    def train(self, buffer, batch_size):  # diffsrsac_agent.py:320-343
        self.steps += 1
        for _ in range(self.K):
            b = buffer.sample(batch_size)
            info = self.feature_step(b)
        info.update(self.critic_step(b))
        info.update(self.update_actor_and_alpha(b))
        self.update_target()
        return info


ORACLES = {"sac": OracleSAC, "ctrlsac": OracleCTRLSAC, "vlsac": OracleVLSAC, "spedersac": OracleSPEDERSAC,
           "diffsrsac": OracleDIFFSRSAC}
