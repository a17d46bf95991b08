"""Soak: long random-action runs at full size; every output finite, no non-converged substeps, no unsupported poses."""
import sys
sys.path.insert(0, ".")
import torch
from balance_robot_b200 import make_vec
for env_id, n, steps in (("Env01-v2", 65536, 3000), ("Env01-v3", 65536, 1500), ("Env01-v1", 1 << 20, 300), ("Env03-v2", 65536, 400)):
    env = make_vec(env_id, n, seed=7)
    obs = env.reset()
    gen = torch.Generator(device="cuda").manual_seed(1)
    bad = torch.zeros((), device="cuda")
    amax = torch.zeros((), device="cuda")
    for k in range(steps):
        obs, r, d, info = env.step(torch.rand((n, 2), device="cuda", generator=gen) * 2 - 1)
        bad += (~torch.isfinite(obs)).sum() + (~torch.isfinite(r)).sum()
        amax = torch.maximum(amax, obs.abs().max())
    q, v, _ = env.get_state()
    st = env.stats()
    print(env_id, n, steps, "non-finite outputs", int(bad.item()), "max |obs|", float(amax.item()), "state finite", bool(torch.isfinite(q).all() and torch.isfinite(v).all()),
          {k: st[k] for k in ("nonconverged", "unsupported", "episodes", "coupled_fallbacks")}, flush=True)
    env.close()
