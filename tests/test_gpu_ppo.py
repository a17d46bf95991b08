"""-m gpu: on-device PPO against the real CUDA envs (rollout buffer, GAE and updates never leave the GPU)."""
import pathlib
import subprocess
import sys

import pytest
import torch

from balance_robot_b200 import make_vec
from balance_robot_b200.ppo import PPO, PPOConfig, evaluate_policy

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_ppo_improves_return_on_env01_v1():
    env = make_vec("Env01-v1", 2048, seed=0)
    agent = PPO(env, PPOConfig(n_steps=32, seed=0), device="cuda:0")
    first = agent.collect_rollouts()
    agent.train()
    for _ in range(24):
        last = agent.collect_rollouts()
        agent.train()
    assert agent.buf["obs"].is_cuda and agent.buf["adv"].is_cuda
    assert last["ep_rew_mean"] > 1.15 * first["ep_rew_mean"], (first, last)
    mean_r, std_r, lens = evaluate_policy(agent.policy, make_vec("Env01-v1", 64, seed=5), 10, True, 6000)
    assert mean_r > first["ep_rew_mean"]
    env.close()


def test_cli_train_writes_reference_style_artifacts(tmp_path):
    cmd = [sys.executable, str(ROOT / "sb_rl.py"), "-a", "PPO", "train", "-e", "Env01-v2", "--num-envs", "1024",
           "--total-timesteps", "200000", "--n-steps", "16"]
    res = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    run = tmp_path / "models" / "Env01-v2_PPO"
    assert (run / "best_model.zip").exists()                                   # EvalCallback target, sb_rl.py:542
    assert list(run.glob("Env01-v2_PPO_cp__*_steps.zip"))                      # CheckpointCallback naming, sb_rl.py:545-550
    assert (tmp_path / "logs").is_dir() and (tmp_path / "movies").is_dir()
    # `test` with the default model name (sb_rl.py:147-149), headless here
    res_t = subprocess.run([sys.executable, str(ROOT / "sb_rl.py"), "-a", "PPO", "test", "-e", "Env01-v2", "--episodes", "6", "--show-io"],
                           cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert res_t.returncode == 0 and "episodes=6 mean_return=" in res_t.stdout, res_t.stderr[-2000:]
    # fine-tune from the saved model with -m on Env03-v2, as README.md:62 does (BASELINE.json configs[3])
    cmd2 = [sys.executable, str(ROOT / "sb_rl.py"), "-a", "PPO", "-m", str(run / "best_model.zip"), "train", "-e", "Env03-v2",
            "--num-envs", "512", "--total-timesteps", "20000", "--n-steps", "16"]
    res2 = subprocess.run(cmd2, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert res2.returncode == 0, res2.stderr[-3000:]
    assert (tmp_path / "models" / "Env03-v2_PPO" / "Env03-v2_PPO_final.zip").exists()


def test_policy_rollout_throughput_path_runs():
    """configs[2]-style rollout: Env01-v3 with on-device policy inference feeding the step kernel (small size here)."""
    env = make_vec("Env01-v3", 8192, seed=1)
    pol = PPO(env, PPOConfig(n_steps=4, seed=1), device="cuda:0").policy
    obs = env.reset()
    for _ in range(10):
        a, _, _ = pol.act(obs)
        obs, r, d, info = env.step(a.clamp(-1, 1))
    assert torch.isfinite(obs).all() and torch.isfinite(r).all()
    env.close()


def test_fused_policy_forward_matches_the_torch_policy():
    """brb_policy_act (one launch: both towers, sample, log-prob, value, clipped action) against the plain PyTorch fp32
    MlpPolicy.act on the same weights, observations and noise draws."""
    from balance_robot_b200.ppo import MlpPolicy
    torch.manual_seed(3)
    pol = MlpPolicy().cuda()
    with torch.no_grad():                                   # move the weights off their orthogonal init
        for p in pol.parameters():
            p.add_(0.05 * torch.randn_like(p))
        pol.log_std.copy_(torch.tensor([-0.3, 0.2]))
    n = 100_003                                             # not a multiple of the CTA size, several grid-stride passes
    obs = torch.randn((n, 6), device="cuda") * torch.tensor([1.5, 3.0, 1.0, 1.0, 1.0, 1.0], device="cuda")
    g1 = torch.Generator(device="cuda").manual_seed(11)
    g2 = torch.Generator(device="cuda").manual_seed(11)
    a_ref, v_ref, lp_ref = pol.act(obs, generator=g1)
    a, v, lp, ac = pol.act_fused(obs, generator=g2)
    assert (a - a_ref).abs().max() < 1e-5 and (v - v_ref).abs().max() < 1e-5
    assert (lp - lp_ref).abs().max() < 1e-4                 # log-prob has a 1 / (2 var) ~ 1 gain on (a - mean)^2 up to ~10
    assert torch.equal(ac, a.clamp(-1.0, 1.0))
    assert (pol.value_fused(obs) - v_ref).abs().max() < 1e-5
    a_det, _, lp_det, _ = pol.act_fused(obs, deterministic=True)
    a_det_ref, _, lp_det_ref = pol.act(obs, deterministic=True)
    assert (a_det - a_det_ref).abs().max() < 1e-5 and (lp_det - lp_det_ref).abs().max() < 1e-5
    # throughput of the fused call at BASELINE configs[2] size (reported, with a loose floor)
    big = torch.randn((1 << 20, 6), device="cuda")
    params = pol.pack_params()
    for _ in range(3):
        pol.act_fused(big, generator=g2, params=params)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        pol.act_fused(big, generator=g2, params=params)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"fused policy forward, 1M robots: {ms:.3f} ms ({(1 << 20) / ms * 1e3:.3e} robots/s)")
    assert ms < 3.0


@pytest.mark.parametrize("path", ["tcgen05", "ffma"])
def test_fused_ppo_gradient_matches_autograd(path, monkeypatch):
    """brb_ppo_grad (forward + clipped-surrogate / value loss + backward, two launches) against PyTorch autograd on the same
    minibatch: gradient of every parameter and the four logged statistics.  path = tcgen05: the tensor-core kernel
    (csrc/brb_policy_tc.cu, bf16 hi/lo x 3 passes, TMEM accumulators) — the default; ffma: the CUDA-core kernel (BRB_PPO_NO_TC)."""
    import ctypes as C
    if path == "ffma":
        monkeypatch.setenv("BRB_PPO_NO_TC", "1")
    else:
        monkeypatch.delenv("BRB_PPO_NO_TC", raising=False)
    from balance_robot_b200 import _cabi
    env = make_vec("Env01-v1", 4096, seed=2)
    cfg = PPOConfig(n_steps=8, seed=2, ent_coef=0.01)
    agent = PPO(env, cfg, device="cuda:0")
    agent.collect_rollouts()
    pol, b = agent.policy, agent.buf
    torch.manual_seed(5)
    with torch.no_grad():                                    # move away from the rollout policy: ratios != 1, some get clipped
        for p in pol.parameters():
            p.add_(0.03 * torch.randn_like(p))
    total = 8 * 4096
    flat = {k: v.reshape(total, *v.shape[2:]) for k, v in b.items()}
    mb = 10_001                                              # not a multiple of the 128-sample tiles
    idx = torch.randperm(total, device="cuda")[:mb].contiguous()
    adv = flat["adv"][idx]
    astats = torch.stack([adv.mean(), 1.0 / (adv.std() + 1e-8)])
    advn = (adv - astats[0]) * astats[1]
    values, logp, entropy = pol.evaluate_actions(flat["obs"][idx], flat["actions"][idx])
    ratio = torch.exp(logp - flat["logp"][idx])
    pl = -torch.min(advn * ratio, advn * torch.clamp(ratio, 1 - cfg.clip_range, 1 + cfg.clip_range)).mean()
    vl = torch.nn.functional.mse_loss(flat["ret"][idx], values)
    loss = pl + cfg.vf_coef * vl - cfg.ent_coef * entropy.mean()
    pol.zero_grad()
    loss.backward()
    g_ref = torch.cat([p.grad.reshape(-1) for p in pol.packed_parameters()])
    lr = (logp - flat["logp"][idx]).detach()
    s_ref = torch.stack([pl.detach(), vl.detach(), ((torch.exp(lr) - 1) - lr).mean(), ((ratio.detach() - 1).abs() > cfg.clip_range).float().mean()])
    assert float(s_ref[3]) > 0.01                            # the clipped branch is exercised

    g = torch.zeros(_cabi.POLICY_NPARAM, device="cuda"); st = torch.zeros(4, device="cuda")
    params = pol.pack_params()
    _cabi.check(_cabi.lib().brb_ppo_grad(params.data_ptr(), flat["obs"].data_ptr(), flat["actions"].data_ptr(), flat["logp"].data_ptr(),
                                         flat["adv"].contiguous().data_ptr(), flat["ret"].contiguous().data_ptr(), idx.data_ptr(), mb,
                                         astats.data_ptr(), cfg.clip_range, cfg.vf_coef, cfg.ent_coef, g.data_ptr(), st.data_ptr(),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)), "brb_ppo_grad")
    torch.cuda.synchronize()
    assert _cabi.lib().brb_ppo_tc_fault(0) == 0              # no pipeline wait of the tcgen05 kernel timed out
    err = (g - g_ref).abs().max() / g_ref.abs().max()
    print(f"{path}: max gradient error / max |gradient| = {float(err):.2e}")
    assert err < 2e-4, float(err)
    off = 0
    for p in pol.packed_parameters():                        # and block by block, relative to each block's own scale
        k = p.numel()
        e = (g[off:off + k] - g_ref[off:off + k]).abs().max() / (g_ref[off:off + k].abs().max() + 1e-12)
        assert e < 1e-3, (off, float(e))
        off += k
    assert torch.allclose(st, s_ref, rtol=1e-3, atol=1e-6), (st, s_ref)
    # throughput at a BASELINE configs[4]-sized minibatch (reported; the tensor-core path must not be slower than the FFMA one)
    big = 1 << 20
    bobs, bact = torch.randn((big, 6), device="cuda"), torch.randn((big, 2), device="cuda")
    blp, badv, bret = torch.randn(big, device="cuda") * 0.1 - 2.0, torch.randn(big, device="cuda"), torch.randn(big, device="cuda")
    bidx = torch.randperm(big, device="cuda")

    def run():
        _cabi.check(_cabi.lib().brb_ppo_grad(params.data_ptr(), bobs.data_ptr(), bact.data_ptr(), blp.data_ptr(), badv.data_ptr(), bret.data_ptr(),
                                             bidx.data_ptr(), big, astats.data_ptr(), cfg.clip_range, cfg.vf_coef, cfg.ent_coef, g.data_ptr(),
                                             st.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "brb_ppo_grad")
    for _ in range(2):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{path}: brb_ppo_grad, 1M samples: {ms:.3f} ms ({big / ms * 1e3:.3e} samples/s)")
    assert _cabi.lib().brb_ppo_tc_fault(0) == 0 and torch.isfinite(g).all()
    env.close()


def test_ppo_learns_to_balance_env01_v3_with_sb3_style_settings():
    """BASELINE configs[0]-style run (`sb_rl.py -a PPO train -e ...`, few envs, long rollouts, SB3's epochs / clip / lr):
    on Env01-v3 the robot goes from falling within ~30 steps to holding the 6,000-step TimeLimit in under a million steps
    (scripts/learn_probe.py: ep_len_mean 6000 after 650k steps, 13 s of wall clock)."""
    env = make_vec("Env01-v3", 64, seed=0)
    agent = PPO(env, PPOConfig(n_steps=256, n_minibatches=32, seed=0), device="cuda:0")
    first = agent.collect_rollouts()
    agent.train()
    best = 0.0
    for _ in range(64):                                     # 64 x 64 x 256 = 1.05M steps
        roll = agent.collect_rollouts()
        agent.train()
        if roll["episodes"] > 0:
            best = max(best, roll["ep_len_mean"])
    assert first["ep_len_mean"] < 100, first
    assert best > 1500, (first, best)
    env.close()


def test_random_permutation_kernel_is_a_keyed_bijection():
    """brb_random_permutation: every value of [0, n) exactly once for sizes on and off powers of two, different keys give different
    orders, and the order looks shuffled (a minibatch of consecutive output positions covers the index range evenly)."""
    import ctypes as C
    from balance_robot_b200 import _cabi
    L = _cabi.lib()
    for n in (1, 2, 3, 1000, 4096, 65536 * 16 + 17, 1 << 24):
        out = torch.empty(n, dtype=torch.int64, device="cuda")
        _cabi.check(L.brb_random_permutation(out.data_ptr(), n, 12345, None), "brb_random_permutation")
        torch.cuda.synchronize()
        assert torch.equal(torch.sort(out).values, torch.arange(n, device="cuda")), n
        if n >= 1000:
            out2 = torch.empty_like(out)
            _cabi.check(L.brb_random_permutation(out2.data_ptr(), n, 12346, None), "brb_random_permutation")
            assert (out != out2).float().mean() > 0.99
            assert (out == torch.arange(n, device="cuda")).float().mean() < 0.01               # few fixed points
            q = out[: n // 4].double()
            assert abs(q.mean().item() / (n - 1) - 0.5) < 0.05 and abs(q.std().item() / (n - 1) - 12 ** -0.5) < 0.03     # uniform over the range
            d = (out[1:] - out[:-1]).double()
            assert abs(torch.corrcoef(torch.stack([out[:-1].double(), out[1:].double()]))[0, 1].item()) < 0.05, n       # neighbours unrelated
