"""Soak of the opt-in wheel-block path: Env03-v2 at full size, random and zero actions; every output finite, nothing non-converged, no unsupported pose."""
import sys
sys.path.insert(0, ".")
import torch
from balance_robot_b200 import make_vec
for policy, steps in (("random", 500), ("zero", 300)):
    n = 65536
    env = make_vec("Env03-v2", n, seed=11, wheel_block=True)
    obs = env.reset()
    gen = torch.Generator(device="cuda").manual_seed(1)
    bad = torch.zeros((), device="cuda"); amax = torch.zeros((), device="cuda")
    for k in range(steps):
        a = torch.rand((n, 2), device="cuda", generator=gen) * 2 - 1 if policy == "random" else torch.zeros((n, 2), device="cuda")
        obs, r, d, info = env.step(a)
        bad += (~torch.isfinite(obs)).sum() + (~torch.isfinite(r)).sum()
        amax = torch.maximum(amax, obs.abs().max())
    q, v, _ = env.get_state()
    st = env.stats()
    print(policy, n, steps, "non-finite outputs", int(bad.item()), "max |obs| %.2f" % float(amax.item()), "state finite", bool(torch.isfinite(q).all() and torch.isfinite(v).all()),
          "max |block v| %.1f" % float(v[:, 8:11].abs().max()), {k: st[k] for k in ("nonconverged", "unsupported", "episodes", "coupled_fallbacks", "coupled_substeps")}, flush=True)
    env.close()
