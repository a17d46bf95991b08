"""On-device PPO for the batched balance-robot environments.

The reference trains with stable_baselines3.PPO("MlpPolicy", env, device='cpu') and SB3's defaults
(sb_rl.py:63-71): MlpPolicy pi=[64,64] vf=[64,64] tanh, state-independent log_std (init 0), orthogonal init
(gains sqrt(2) / 0.01 / 1), n_steps 2048, batch 64, 10 epochs, lr 3e-4, gamma 0.99, lambda 0.95, clip 0.2,
vf_coef 0.5, ent_coef 0, max_grad_norm 0.5, Adam eps 1e-5, per-minibatch advantage normalisation, actions
clipped to the action space for the env while the buffer keeps the unclipped sample, and gamma * V(terminal
observation) added to the reward of time-limit truncations.  This module keeps those semantics and
hyper-parameters but runs the rollout buffer, GAE and the updates on the GPU next to the env kernels, with
rollout / batch sizes that scale with the number of envs (SB3's 2048 x 1 env is for a single CPU env).

Multi-GPU: one process per GPU (torchrun), envs sharded by rank; the only collectives are one flat-gradient
all-reduce per minibatch and one all-reduce of the rollout statistics (NCCL over NVLink; gloo in CPU tests).
Checkpoints are written in Stable-Baselines3's model.zip layout (sb3_format.py: `data` with SB3's attribute names, the
policy class and the Box spaces as version-independent pickle streams, SB3's state-dict key names); zips written by SB3's
PPO("MlpPolicy") load here.  Loading OUR zips into real SB3 could not be run in this image (SB3 is not installable).
"""
from __future__ import annotations

import dataclasses
import io
import json
import math
import pathlib
import time
import zipfile
from typing import Callable, Dict, Optional

import torch
import torch.nn as nn


# ------------------------------------------------------------------------------------------------ policy
class _MlpExtractor(nn.Module):
    def __init__(self, obs_dim: int, hidden: int):
        super().__init__()
        self.policy_net = nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh())
        self.value_net = nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh())


class MlpPolicy(nn.Module):
    """SB3 ActorCriticPolicy("MlpPolicy") for Box observations / actions; same parameter names and init."""

    def __init__(self, obs_dim: int = 6, act_dim: int = 2, hidden: int = 64, log_std_init: float = 0.0):
        super().__init__()
        self.mlp_extractor = _MlpExtractor(obs_dim, hidden)
        self.action_net = nn.Linear(hidden, act_dim)
        self.value_net = nn.Linear(hidden, 1)
        self.log_std = nn.Parameter(torch.full((act_dim,), float(log_std_init)))
        for mod, gain in ((self.mlp_extractor, math.sqrt(2)), (self.action_net, 0.01), (self.value_net, 1.0)):
            for m in mod.modules():
                if isinstance(m, nn.Linear):
                    nn.init.orthogonal_(m.weight, gain=gain)
                    nn.init.zeros_(m.bias)

    def _dist(self, obs):
        mean = self.action_net(self.mlp_extractor.policy_net(obs))
        return mean, self.log_std.expand_as(mean)

    def value(self, obs):
        return self.value_net(self.mlp_extractor.value_net(obs)).squeeze(-1)

    @staticmethod
    def _log_prob(actions, mean, log_std):
        var = torch.exp(2 * log_std)
        return (-((actions - mean) ** 2) / (2 * var) - log_std - 0.5 * math.log(2 * math.pi)).sum(-1)

    def packed_parameters(self):
        """The parameters in the order of the flat block the CUDA kernels read (include/brb.h: SB3 state-dict order)."""
        return (self.mlp_extractor.policy_net[0].weight, self.mlp_extractor.policy_net[0].bias,
                self.mlp_extractor.policy_net[2].weight, self.mlp_extractor.policy_net[2].bias,
                self.mlp_extractor.value_net[0].weight, self.mlp_extractor.value_net[0].bias,
                self.mlp_extractor.value_net[2].weight, self.mlp_extractor.value_net[2].bias,
                self.action_net.weight, self.action_net.bias, self.value_net.weight, self.value_net.bias, self.log_std)

    def pack_params(self) -> torch.Tensor:
        """Flat fp32 parameter block = the layout brb_policy_act / brb_ppo_grad read."""
        return torch.cat([p.detach().reshape(-1) for p in self.packed_parameters()]).float().contiguous()

    @torch.no_grad()
    def act_fused(self, obs, deterministic: bool = False, generator: Optional[torch.Generator] = None, params: Optional[torch.Tensor] = None):
        """act() as ONE CUDA launch (csrc/brb_policy.cu): both towers, the Gaussian sample, log-prob, value and the clipped
        copy of the action for the env.  Returns (actions, values, log_prob, clipped_actions).  `params` = pack_params()
        (pass it when the weights do not change between calls, as during a rollout)."""
        import ctypes as C
        from . import _cabi
        if not obs.is_cuda:
            raise _cabi.BrbError("act_fused runs on CUDA tensors only")
        if self.action_net.in_features != 64 or self.action_net.out_features != 2 or obs.shape[-1] != 6:
            raise _cabi.BrbError("act_fused is built for the 6-64-64-2 MlpPolicy")
        params = self.pack_params() if params is None else params
        assert params.numel() == _cabi.POLICY_NPARAM and params.is_cuda
        obs = obs.to(torch.float32).contiguous()
        n = obs.shape[0]
        noise = None if deterministic else torch.randn((n, 2), device=obs.device, dtype=torch.float32, generator=generator)
        actions = torch.empty((n, 2), device=obs.device); clipped = torch.empty((n, 2), device=obs.device)
        values = torch.empty(n, device=obs.device); logp = torch.empty(n, device=obs.device)
        with torch.cuda.device(obs.device):
            _cabi.check(_cabi.lib().brb_policy_act(params.data_ptr(), obs.data_ptr(), noise.data_ptr() if noise is not None else None, n,
                                                   actions.data_ptr(), clipped.data_ptr(), values.data_ptr(), logp.data_ptr(),
                                                   C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)), "brb_policy_act")
        return actions, values, logp, clipped

    @torch.no_grad()
    def value_fused(self, obs, params: Optional[torch.Tensor] = None):
        """value() through the fused kernel (critic tower only)."""
        import ctypes as C
        from . import _cabi
        params = self.pack_params() if params is None else params
        obs = obs.to(torch.float32).contiguous()
        values = torch.empty(obs.shape[0], device=obs.device)
        with torch.cuda.device(obs.device):
            _cabi.check(_cabi.lib().brb_policy_act(params.data_ptr(), obs.data_ptr(), None, obs.shape[0], None, None, values.data_ptr(), None,
                                                   C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)), "brb_policy_act")
        return values

    @torch.no_grad()
    def value_masked_fused(self, obs, mask, params: torch.Tensor):
        """critic values of the rows where mask (uint8) != 0, zero elsewhere (csrc/brb_policy.cu: brb_policy_value_masked)."""
        import ctypes as C
        from . import _cabi
        obs = obs.to(torch.float32).contiguous()
        mask = mask.to(torch.uint8).contiguous()
        values = torch.empty(obs.shape[0], device=obs.device)
        with torch.cuda.device(obs.device):
            _cabi.check(_cabi.lib().brb_policy_value_masked(params.data_ptr(), obs.data_ptr(), mask.data_ptr(), obs.shape[0], values.data_ptr(),
                                                            C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)), "brb_policy_value_masked")
        return values

    @torch.no_grad()
    def act(self, obs, deterministic: bool = False, generator: Optional[torch.Generator] = None):
        mean, log_std = self._dist(obs)
        if deterministic:
            actions = mean
        else:
            noise = torch.randn(mean.shape, device=mean.device, dtype=mean.dtype, generator=generator)
            actions = mean + noise * torch.exp(log_std)
        return actions, self.value(obs), self._log_prob(actions, mean, log_std)

    def evaluate_actions(self, obs, actions):
        mean, log_std = self._dist(obs)
        entropy = (0.5 + 0.5 * math.log(2 * math.pi) + log_std).sum(-1)
        return self.value(obs), self._log_prob(actions, mean, log_std), entropy

    def predict(self, obs, deterministic: bool = True):
        """SB3-style predict(): returns (clipped actions, None)."""
        a, _, _ = self.act(torch.as_tensor(obs, device=self.log_std.device, dtype=torch.float32), deterministic)
        return a.clamp(-1.0, 1.0), None


# ------------------------------------------------------------------------------------------------ GAE
def compute_gae(rewards, values, dones, last_values, gamma: float, lam: float):
    """SB3 RolloutBuffer.compute_returns_and_advantage.  rewards/values/dones: [T, N]; dones[t] = episode ended at t
    (so the value of step t+1 belongs to a new episode)."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    last = torch.zeros_like(last_values)
    for t in reversed(range(T)):
        next_values = last_values if t == T - 1 else values[t + 1]
        nonterminal = 1.0 - dones[t]
        delta = rewards[t] + gamma * next_values * nonterminal - values[t]
        last = delta + gamma * lam * nonterminal * last
        adv[t] = last
    return adv, adv + values


# ------------------------------------------------------------------------------------------------ trainer
@dataclasses.dataclass
class PPOConfig:
    n_steps: int = 32               # rollout length per env (SB3: 2048 for ONE env)
    n_epochs: int = 10
    n_minibatches: int = 4          # SB3: batch_size 64 -> 32 minibatches; here the count is fixed instead
    learning_rate: float = 3e-4
    gamma: float = 0.99
    gae_lambda: float = 0.95
    clip_range: float = 0.2
    vf_coef: float = 0.5
    ent_coef: float = 0.0
    max_grad_norm: float = 0.5
    adam_eps: float = 1e-5
    normalize_advantage: bool = True
    seed: int = 0


class PPO:
    def __init__(self, env, config: PPOConfig = PPOConfig(), policy: Optional[MlpPolicy] = None, device=None,
                 rank: int = 0, world_size: int = 1):
        self.env, self.cfg, self.rank, self.world = env, config, rank, world_size
        self.device = torch.device(device if device is not None else getattr(env, "device", "cpu"))
        torch.manual_seed(config.seed)          # identical initial weights on every rank
        self.policy = (policy or MlpPolicy()).to(self.device)
        self.optimizer = torch.optim.Adam(self.policy.parameters(), lr=config.learning_rate, eps=config.adam_eps)
        self.gen = torch.Generator(device=self.device).manual_seed(config.seed * 1000003 + rank)
        self.num_timesteps = 0
        self.fused_update = True      # CUDA: minibatch forward + loss + backward through brb_ppo_grad (False: autograd)
        self._gflat = self._pflat = None
        self._adam_t = 0
        self._comm = None
        if self.device.type == "cuda" and self.policy.action_net.in_features == 64:
            self._flatten_parameters()
            if world_size > 1:
                self._setup_peer_comm()
        self._obs = None
        self.ep_stats = {"return_sum": 0.0, "len_sum": 0.0, "count": 0.0}
        n, T = env.num_envs, config.n_steps
        dv = self.device
        self.buf = {"obs": torch.zeros((T, n, 6), device=dv), "actions": torch.zeros((T, n, 2), device=dv),
                    "logp": torch.zeros((T, n), device=dv), "values": torch.zeros((T, n), device=dv),
                    "rewards": torch.zeros((T, n), device=dv), "dones": torch.zeros((T, n), device=dv)}

    # ---- flat parameter / gradient / Adam-moment blocks for the fused CUDA update (csrc/brb_policy.cu)
    def _flatten_parameters(self) -> None:
        """Re-homes every parameter (and its .grad) as a view of ONE flat fp32 block in the layout brb_policy_act /
        brb_ppo_grad / brb_adam_clip_step read, so no packing, no per-parameter launches and no copies are needed."""
        from . import _cabi
        ps = self.policy.packed_parameters()
        flat = torch.cat([p.detach().reshape(-1) for p in ps]).float().contiguous()
        assert flat.numel() == _cabi.POLICY_NPARAM
        self._pflat, self._gflat = flat, torch.zeros_like(flat)
        self._m, self._v = torch.zeros_like(flat), torch.zeros_like(flat)
        self._gstats = torch.zeros(4, device=self.device)
        self._gnorm = torch.zeros(1, device=self.device)
        off = 0
        for p in ps:
            k = p.numel()
            p.data = flat[off:off + k].view_as(p)
            p.grad = self._gflat[off:off + k].view_as(p)
            off += k

    def _setup_peer_comm(self) -> None:
        """Peer-memory block for the fused all-reduce + clip + Adam kernel (csrc/brb_policy.cu: brb_comm_*): every rank exports its
        block's cudaIpc handle, the handles are all-gathered through torch.distributed, every rank maps its peers' blocks.
        Falls back to the NCCL all-reduce (+ brb_adam_clip_step) if any rank cannot map a peer, or with BRB_PPO_NO_P2P=1."""
        import ctypes as C
        import os
        import torch.distributed as dist
        from . import _cabi
        if os.environ.get("BRB_PPO_NO_P2P") or not dist.is_initialized() or self.world > 8:
            return
        L = _cabi.lib()
        comm = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        rc = L.brb_comm_create(self.rank, self.world, dev_index, self._pflat.numel(), C.byref(comm))
        handle = (C.c_ubyte * 64)()
        if rc == 0:
            rc = L.brb_comm_export(comm, handle)
        mine = torch.tensor(list(bytes(handle)) + [1 if rc == 0 else 0], dtype=torch.uint8, device=self.device)
        gathered = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine)
        rows = [bytes(g.cpu().tolist()) for g in gathered]
        ok = all(r[64] == 1 for r in rows)
        if ok:
            ok = L.brb_comm_open(comm, b"".join(r[:64] for r in rows)) == 0
        flag = torch.tensor([1.0 if ok else 0.0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if float(flag.item()) == 1.0:
            self._comm = comm
        elif comm:
            L.brb_comm_destroy(comm)

    def close(self) -> None:
        if self._comm is not None:
            from . import _cabi
            _cabi.lib().brb_comm_destroy(self._comm)
            self._comm = None

    def _export_adam_state(self) -> None:
        """flat Adam moments -> the torch optimizer's per-parameter state (for checkpoints in SB3's layout)."""
        if self._pflat is None or self._adam_t == 0:
            return
        off = 0
        for p in self.policy.packed_parameters():
            k = p.numel()
            self.optimizer.state[p] = {"step": torch.tensor(float(self._adam_t)), "exp_avg": self._m[off:off + k].view_as(p).clone(),
                                       "exp_avg_sq": self._v[off:off + k].view_as(p).clone()}
            off += k

    def _import_adam_state(self) -> None:
        if self._pflat is None:
            return
        off = 0
        for p in self.policy.packed_parameters():
            k = p.numel()
            st = self.optimizer.state.get(p)
            if st and "exp_avg" in st:
                self._m[off:off + k].copy_(st["exp_avg"].reshape(-1))
                self._v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                self._adam_t = int(float(st["step"]))
            off += k

    # ---- distributed helpers (no-ops for world_size 1)
    def _all_reduce_(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def _sync_grads(self) -> None:
        if self.world == 1:
            return
        params = [p for p in self.policy.parameters() if p.grad is not None]
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        self._all_reduce_(flat).div_(self.world)
        off = 0
        for p in params:
            k = p.numel()
            p.grad.copy_(flat[off:off + k].view_as(p))
            off += k

    # ---- rollout
    def collect_rollouts(self) -> Dict[str, float]:
        cfg, env, b = self.cfg, self.env, self.buf
        if self._obs is None:
            self._obs = torch.as_tensor(env.reset(), device=self.device, dtype=torch.float32).clone()
        ep_ret = torch.zeros((), device=self.device, dtype=torch.float64)
        ep_len = torch.zeros((), device=self.device, dtype=torch.float64)
        ep_cnt = torch.zeros((), device=self.device, dtype=torch.float64)
        fused = self.device.type == "cuda" and self.policy.action_net.in_features == 64
        params = (self._pflat if self._pflat is not None else self.policy.pack_params()) if fused else None   # the weights do not change during a rollout
        for t in range(cfg.n_steps):
            if fused:       # one launch: towers + sample + log-prob + value + clipped copy (csrc/brb_policy.cu)
                actions, values, logp, clipped = self.policy.act_fused(self._obs, generator=self.gen, params=params)
            else:
                actions, values, logp = self.policy.act(self._obs, generator=self.gen)
                clipped = actions.clamp(-1.0, 1.0)                     # SB3 clips for the env only
            b["obs"][t].copy_(self._obs)
            b["actions"][t].copy_(actions)
            b["values"][t].copy_(values)
            b["logp"][t].copy_(logp)
            obs, rew, done, infos = env.step(clipped)
            done_f = done.to(torch.float32)
            rew = rew.clone()
            # TimeLimit bootstrap: reward += gamma * V(terminal_observation) where the episode was truncated
            # (a `trunc.any()` test would put a host sync into every step of the rollout; the CUDA path evaluates the critic only for the
            # flagged rows on the device, <= N / max_episode_steps of them per step; the CPU path evaluates every row and masks)
            if hasattr(infos, "truncated"):
                with torch.no_grad():
                    if fused:
                        tv = self.policy.value_masked_fused(infos.terminal_observation, infos.truncated, params)
                    else:
                        tv = self.policy.value(infos.terminal_observation) * infos.truncated.to(rew.dtype)
                rew = rew + cfg.gamma * tv
            b["rewards"][t].copy_(rew)
            b["dones"][t].copy_(done_f)
            if hasattr(infos, "episode_return"):
                d64 = done_f.double()
                ep_ret += (infos.episode_return.double() * d64).sum()
                ep_len += (infos.episode_length.double() * d64).sum()
                ep_cnt += d64.sum()
            self._obs = obs.clone()
        with torch.no_grad():
            last_values = self.policy.value_fused(self._obs, params) if fused else self.policy.value(self._obs)
        adv, ret = compute_gae(b["rewards"], b["values"], b["dones"], last_values, cfg.gamma, cfg.gae_lambda)
        b["adv"], b["ret"] = adv, ret
        self.num_timesteps += cfg.n_steps * env.num_envs * self.world
        stats = self._all_reduce_(torch.stack([ep_ret, ep_len, ep_cnt]))
        r, l, c = (float(x) for x in stats.tolist())
        self.ep_stats = {"return_sum": r, "len_sum": l, "count": c}
        return {"ep_rew_mean": r / c if c else float("nan"), "ep_len_mean": l / c if c else float("nan"), "episodes": c}

    # ---- update
    def _train_fused(self) -> Dict[str, float]:
        """train() as three CUDA launches per minibatch (csrc/brb_policy.cu): brb_ppo_grad (forward + loss + backward, one launch per
        tower) and brb_adam_clip_step (gradient averaging, global-norm clipping and Adam on the flat parameter block); across
        ranks one all-reduce of the flat gradient in between."""
        import ctypes as C
        from . import _cabi
        cfg, b = self.cfg, self.buf
        T, n = b["rewards"].shape
        total = T * n
        mb = total // cfg.n_minibatches
        obs, act = b["obs"].reshape(total, 6), b["actions"].reshape(total, 2)
        oldlp, adv, ret = b["logp"].reshape(total), b["adv"].reshape(total).contiguous(), b["ret"].reshape(total).contiguous()
        L, stream = _cabi.lib(), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        unit = torch.tensor([0.0, 1.0], device=self.device)
        lr, eps = self.optimizer.param_groups[0]["lr"], self.optimizer.param_groups[0]["eps"]
        b1, b2 = self.optimizer.param_groups[0]["betas"]
        self._gflat.zero_(); self._gstats.zero_()
        updates = 0
        if getattr(self, "_perm", None) is None or self._perm.numel() != total:
            self._perm = torch.empty(total, dtype=torch.int64, device=self.device)
        perm = self._perm
        for _ in range(cfg.n_epochs):
            # minibatch order of this epoch: keyed bijection of [0, total) written by one kernel (csrc/brb_policy.cu), no sort
            self._perm_calls = getattr(self, "_perm_calls", 0) + 1
            with torch.cuda.device(self.device):
                _cabi.check(L.brb_random_permutation(perm.data_ptr(), total, (int(cfg.seed) * 1000003 + self._perm_calls) & (2 ** 64 - 1), stream),
                            "brb_random_permutation")
            for k in range(cfg.n_minibatches):
                idx = perm[k * mb:(k + 1) * mb]
                if cfg.normalize_advantage and mb > 1:
                    a = adv[idx]
                    sd, mu = torch.std_mean(a)
                    astats = torch.stack([mu, 1.0 / (sd + 1e-8)])
                else:
                    astats = unit
                with torch.cuda.device(self.device):
                    self._adam_t += 1
                    gptr = L.brb_comm_grad(self._comm, self._adam_t) if self._comm is not None else self._gflat.data_ptr()
                    _cabi.check(L.brb_ppo_grad(self._pflat.data_ptr(), obs.data_ptr(), act.data_ptr(), oldlp.data_ptr(), adv.data_ptr(),
                                               ret.data_ptr(), idx.data_ptr(), mb, astats.data_ptr(), cfg.clip_range, cfg.vf_coef,
                                               cfg.ent_coef, gptr, self._gstats.data_ptr(), stream), "brb_ppo_grad")
                    if self._comm is not None:
                        # all-reduce over NVLink peer memory + clipping + Adam: one launch on this stream, no NCCL call
                        _cabi.check(L.brb_comm_allreduce_adam(self._comm, self._pflat.data_ptr(), self._m.data_ptr(), self._v.data_ptr(), lr, b1, b2,
                                                              eps, self._adam_t, cfg.max_grad_norm or 0.0, self._gnorm.data_ptr(), stream),
                                    "brb_comm_allreduce_adam")
                    else:
                        if self.world > 1:
                            self._all_reduce_(self._gflat)
                        _cabi.check(L.brb_adam_clip_step(self._pflat.data_ptr(), self._gflat.data_ptr(), self._m.data_ptr(), self._v.data_ptr(),
                                                         self._pflat.numel(), lr, b1, b2, eps, self._adam_t, cfg.max_grad_norm or 0.0,
                                                         1.0 / self.world, self._gnorm.data_ptr(), stream), "brb_adam_clip_step")
                updates += 1
        vals = (self._gstats / max(1, updates)).tolist()
        return dict(zip(("policy_loss", "value_loss", "approx_kl", "clip_fraction"), vals))

    def train(self) -> Dict[str, float]:
        if self.fused_update and self.device.type == "cuda" and self.policy.action_net.in_features == 64:
            return self._train_fused()
        cfg, b = self.cfg, self.buf
        T, n = b["rewards"].shape
        flat = {k: v.reshape(T * n, *v.shape[2:]) for k, v in b.items()}
        total = T * n
        mb = total // cfg.n_minibatches
        logs = torch.zeros(4, device=self.device)     # policy_loss, value_loss, approx_kl, clip_fraction (summed on device)
        updates = 0
        for _ in range(cfg.n_epochs):
            perm = torch.randperm(total, device=self.device, generator=self.gen)
            for k in range(cfg.n_minibatches):
                idx = perm[k * mb:(k + 1) * mb]
                adv = flat["adv"][idx]
                if cfg.normalize_advantage and adv.numel() > 1:
                    adv = (adv - adv.mean()) / (adv.std() + 1e-8)
                values, logp, entropy = self.policy.evaluate_actions(flat["obs"][idx], flat["actions"][idx])
                ratio = torch.exp(logp - flat["logp"][idx])
                pl = -torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - cfg.clip_range, 1 + cfg.clip_range)).mean()
                vl = torch.nn.functional.mse_loss(flat["ret"][idx], values)
                loss = pl + cfg.vf_coef * vl - cfg.ent_coef * entropy.mean()
                self.optimizer.zero_grad(set_to_none=False)
                loss.backward()
                self._sync_grads()
                torch.nn.utils.clip_grad_norm_(self.policy.parameters(), cfg.max_grad_norm)
                self.optimizer.step()
                with torch.no_grad():
                    lr_ = logp - flat["logp"][idx]
                    logs += torch.stack([pl.detach(), vl.detach(), ((torch.exp(lr_) - 1) - lr_).mean(),
                                         ((ratio - 1).abs() > cfg.clip_range).float().mean()])
                updates += 1
        vals = (logs / max(1, updates)).tolist()
        return dict(zip(("policy_loss", "value_loss", "approx_kl", "clip_fraction"), vals))

    def learn(self, total_timesteps: int, callback: Optional[Callable[["PPO", Dict[str, float]], bool]] = None,
              log_interval: int = 1, writer=None, verbose: int = 1) -> "PPO":
        it, t0 = 0, time.time()
        while self.num_timesteps < total_timesteps:
            roll = self.collect_rollouts()
            upd = self.train()
            it += 1
            self._check_unsupported_poses()
            fps = self.num_timesteps / max(1e-9, time.time() - t0)
            rec = dict(roll, **upd, fps=fps, total_timesteps=self.num_timesteps, iterations=it)
            if self.rank == 0:
                if writer is not None:        # TensorBoard scalar names follow SB3
                    writer.add_scalar("rollout/ep_rew_mean", rec["ep_rew_mean"], self.num_timesteps)
                    writer.add_scalar("rollout/ep_len_mean", rec["ep_len_mean"], self.num_timesteps)
                    writer.add_scalar("time/fps", fps, self.num_timesteps)
                    writer.add_scalar("train/approx_kl", rec["approx_kl"], self.num_timesteps)
                    writer.add_scalar("train/value_loss", rec["value_loss"], self.num_timesteps)
                if verbose and it % log_interval == 0:
                    print(f"[ppo] it {it:4d} steps {self.num_timesteps:>12,d} fps {fps:,.0f} ep_rew_mean {rec['ep_rew_mean']:.2f} "
                          f"ep_len_mean {rec['ep_len_mean']:.1f} kl {rec['approx_kl']:.4f} vloss {rec['value_loss']:.3f}", flush=True)
            if callback is not None and callback(self, rec) is False:
                break
        return self

    def _check_unsupported_poses(self) -> None:
        """The step kernels count env-steps that ended in a pose whose contacts they do not model (chassis on the floor, a
        wheel lying flat): those transitions are simulated without that contact.  Unreachable before the 50 degree
        termination in Env01 (0 in every soak so far), but a trainer must not consume them silently."""
        stats = getattr(self.env, "stats", None)
        if stats is None:
            return
        u = int(stats().get("unsupported", 0))
        if u > getattr(self, "_unsupported_seen", 0):
            import warnings
            warnings.warn(f"{u - getattr(self, '_unsupported_seen', 0)} env-steps ended in a pose with unmodelled contacts "
                          f"(chassis-floor / wheel lying flat; {u} in total): those transitions lack that contact force", RuntimeWarning)
            self._unsupported_seen = u

    # ---- checkpoints: Stable-Baselines3's model.zip (sb3_format.py: data / policy.pth / policy.optimizer.pth / ...)
    def save(self, path) -> pathlib.Path:
        from . import sb3_format
        path = pathlib.Path(path)
        if path.suffix != ".zip":
            path = path.with_suffix(".zip")
        path.parent.mkdir(parents=True, exist_ok=True)
        n_envs = int(getattr(self.env, "num_envs", 1)) * self.world
        osp, asp = getattr(self.env, "observation_space", None), getattr(self.env, "action_space", None)
        obs_low, obs_high = (osp.low.tolist(), osp.high.tolist()) if osp is not None else ([-1.0] * 6, [1.0] * 6)
        act_low, act_high = (asp.low.tolist(), asp.high.tolist()) if asp is not None else ([-1.0] * 2, [1.0] * 2)
        batch = max(1, self.cfg.n_steps * n_envs // self.cfg.n_minibatches)
        data = sb3_format.build_data(self.cfg, self.num_timesteps, n_envs, self._adam_t or 0, obs_low, obs_high, act_low, act_high, batch)

        def blob(obj):
            bio = io.BytesIO()
            torch.save(obj, bio)
            return bio.getvalue()
        self._export_adam_state()
        # SB3 orders the optimizer's params like policy.parameters(); ours is the same module tree, so the state dict lines up
        with zipfile.ZipFile(path, "w") as z:
            z.writestr("data", data)
            z.writestr("pytorch_variables.pth", blob(None))
            z.writestr("policy.pth", blob({k: v.detach().cpu().clone() for k, v in self.policy.state_dict().items()}))
            z.writestr("policy.optimizer.pth", blob(self.optimizer.state_dict()))
            z.writestr("_stable_baselines3_version", sb3_format.SB3_VERSION)
            z.writestr("system_info.txt", sb3_format.SYSTEM_INFO)
        return path

    @classmethod
    def load(cls, path, env, config: Optional[PPOConfig] = None, **kw) -> "PPO":
        """Loads a model.zip written by this package or by Stable-Baselines3's PPO("MlpPolicy") with the default 64-64 tanh
        architecture (algorithm_class.load(model_file, env=env), reference sb_rl.py:519-525)."""
        from . import sb3_format
        path = pathlib.Path(path)
        if not path.exists():
            raise RuntimeError(f"model file {path} does not exist")        # mirrors sb_rl.py:100-101
        with zipfile.ZipFile(path) as z:
            data = sb3_format.parse_data(z.read("data").decode())
            sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=False)
            opt = (torch.load(io.BytesIO(z.read("policy.optimizer.pth")), map_location="cpu", weights_only=False)
                   if "policy.optimizer.pth" in z.namelist() else None)
        fields = sb3_format.ppo_config_fields(data)
        cfg = config or PPOConfig(**{k: v for k, v in fields.items() if k in PPOConfig.__dataclass_fields__})
        self = cls(env, cfg, **kw)
        self.policy.load_state_dict(sd)
        if opt:
            try:
                self.optimizer.load_state_dict(opt)
                self._import_adam_state()
            except Exception:
                pass
        self.num_timesteps = int(data.get("num_timesteps", 0))
        return self


@torch.no_grad()
def evaluate_policy(policy: MlpPolicy, env, n_eval_episodes: int = 5, deterministic: bool = True, max_steps: int = 6000):
    """SB3 evaluate_policy on a (separate) vectorised env: mean / std of the FIRST episode of each of the first
    `n_eval_episodes` envs.  (Taking the first episodes to finish anywhere in the batch would favour short ones —
    12.8 % of Env01-v2 episodes end at their first step, SURVEY.md Q3 — which is why SB3 fixes a per-env quota too.)"""
    n = min(n_eval_episodes, env.num_envs)
    obs = torch.as_tensor(env.reset(), dtype=torch.float32, device=policy.log_std.device)
    ret = torch.zeros(n, device=obs.device)
    length = torch.zeros(n, device=obs.device)
    finished = torch.zeros(n, dtype=torch.bool, device=obs.device)
    for _ in range(max_steps + 1):
        a, _, _ = policy.act(obs, deterministic=deterministic)
        obs, rew, done, infos = env.step(a.clamp(-1.0, 1.0))
        newly = done[:n].to(torch.bool) & ~finished
        if bool(newly.any()):
            ret[newly] = infos.episode_return[:n][newly].to(ret.dtype)
            length[newly] = infos.episode_length[:n][newly].to(length.dtype)
            finished |= newly
        if bool(finished.all()):
            break
    r = ret[finished] if bool(finished.any()) else torch.tensor([float("nan")])
    return float(r.mean()), float(r.std(unbiased=False)), [int(x) for x in length[finished].tolist()]
