#!/usr/bin/env python
"""bench.py — env-steps/s of the fused Env01-v2 step (BASELINE.json metric) on N B200s, device-timed.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference step (oracle), host cores

One "step" = one VecEnv.step over every env of the shard = 250 physics substeps per env + task logic +
auto-reset.  Workload at N=1: BASELINE.json configs[1] — Env01-v2, 65,536 envs, random
actions a ~ U(-1,1)^2 (torch.Generator(device).manual_seed(1234)), Philox noise, auto-reset on.  N>1: the same
shard on every rank (weak scaling), Philox streams keyed by global env id, no collective in the step.

Prints ONE JSON line (rank 0).  Keys beyond the base contract: roofline (FP32-pipe bound: the schema's
"hbm|tensor" does not apply to this path), cpu_baseline, e2e, clocks, gpu_launches.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "env-steps/sec (Env01-v2, device-timed)"
UNIT = "env-steps/s"
ENV_ID = "Env01-v2"
FLOP_PER_SUBSTEP_CONTACT = 3700.0   # 4 contacts, one solver pass, nv = 8
FLOP_PER_SUBSTEP_AIR = 1000.0
FRAME_SKIP = 250
ALG_FLOP_PER_ENV_STEP = FLOP_PER_SUBSTEP_CONTACT * FRAME_SKIP          # 9.25e5 ("9.3e5" in SURVEY.md)
ALG_BYTES_PER_ENV_STEP = 290.0
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12
# Counters of brb_step_kernel<Env01_v2> from the committed `ncu --set full` capture of this command (one launch, 65,536 robots):
# profiles/<round>_step_kernel_ncu_summary.json is written by scripts/make_profile_summaries.py from the raw page next to it.
def load_ncu_summary():
    for name in ("r2_step_kernel_ncu_summary.json", "r1_step_kernel_ncu_summary.json"):
        f = ROOT / "profiles" / name
        if f.exists():
            d = json.load(open(f))
            d["source"] = f"profiles/{name}"
            return d
    return None


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--spinup", type=int, default=192,
                    help="untimed steps before the warm-up: all robots start 2 cm in the air (Q11), so the first ~13 steps after "
                         "reset_all are cheap free fall; the spin-up brings the population to its stationary airborne/grounded mix")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=65536)
    ap.add_argument("--env", default=ENV_ID)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-threads", type=int, default=0, help="threads for the CPU legs (0 = all cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--actions", default="random", choices=["random", "policy"],
                    help="random: U(-1,1)^2 resident in HBM (BASELINE configs[1]); policy: on-device MlpPolicy inference every step")
    ap.add_argument("--sustained-s", type=float, default=1.2, help="length of the back-to-back `sustained` sub-run (0 = skip)")
    ap.add_argument("--ppo-envs-per-gpu", type=int, default=1048576, help="envs per GPU of the `ppo` sub-record (BASELINE configs[4]; 0 = skip)")
    ap.add_argument("--ppo-iters", type=int, default=2)
    return ap.parse_args()


def workload_config(args, world):
    return {"workload": f"{args.env} random-action rollout, {args.envs_per_gpu} envs per GPU x {world} GPU(s), "
                        "250 substeps/step (h=2e-5, implicitfast), auto-reset, Philox noise",
            "env": args.env, "envs_per_gpu": args.envs_per_gpu, "n_envs": args.envs_per_gpu * world,
            "frame_skip": FRAME_SKIP,
            "actions": "U(-1,1)^2, torch.Generator(cuda).manual_seed(1234)" if args.actions == "random"
            else "on-device MlpPolicy (6-64-64-2 tanh, random init seed 0) sampled every step inside the timed region",
            "cache": "L2 flushed (256 MiB write) before every timed step", "sharding": f"env-dp{world}",
            "spinup_steps": args.spinup}


# ------------------------------------------------------------------------------------------------ CPU legs (oracle)
def cpu_leg(args, steps, warmup, threads):
    """Times the fp64 oracle (CPU restatement of the reference step; MuJoCo itself is not installable here)
    over 64 envs per thread with the same action distribution and replayed Philox draws."""
    import numpy as np
    from balance_robot_b200 import mjcf, registry
    from oracle import ref
    spec = registry.spec(args.env)
    n = 64 * threads
    rv = ref.RefVecEnv(mjcf.parse(spec.scene), args.env, n, spec.max_episode_steps, nthreads=threads)
    _, ur = ref.philox_draws(args.seed, 0, n, 0)
    rv.reset(ur)
    rng = np.random.default_rng(1234)
    warmup += args.spinup          # same untimed spin-up to the stationary airborne/grounded mix as the GPU arm
    draws = [ref.philox_draws(args.seed, 0, n, k + 1) for k in range(warmup + steps)]
    acts = rng.uniform(-1, 1, (warmup + steps, n, 2)).astype(np.float32)
    for k in range(warmup):
        rv.step(acts[k], *draws[k])
    t0 = time.perf_counter()
    for k in range(warmup, warmup + steps):
        rv.step(acts[k], *draws[k])
    dt = time.perf_counter() - t0
    rv.close()
    return n * steps / dt, dt, n


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = args.cpu_threads or (os.cpu_count() or 1)
    value, dt, n = cpu_leg(args, args.steps, args.warmup, threads)
    sample = f"{n} envs (64 per thread) x {args.steps} steps after {args.spinup} spin-up + {args.warmup} warm-up, same action distribution, replayed Philox draws"
    line = {"metric": METRIC.replace(ENV_ID, args.env), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, world), "impl": "reference",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference arm = fp64 C restatement of the reference step (oracle/); mujoco/gymnasium/SB3 are not installable offline"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm_sorted = sorted(sm)
        # median over the busier half of the samples (the sampler also sees the idle edges of the region)
        busy = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw)}


def run_b200(args, rank, world, local_rank):
    import torch
    import ctypes as C
    from balance_robot_b200 import make_vec, _cabi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    n = args.envs_per_gpu
    env = make_vec(args.env, n, device=dev, seed=args.seed, env_id_offset=rank * n)
    env.reset()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    nact = 16
    acts = [torch.rand((n, 2), device=dev, generator=gen) * 2 - 1 for _ in range(nact)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    # FP32-pipe peak probe (roofline denominator, measured in this run)
    fl, ms = C.c_double(), C.c_double()
    _cabi.check(_cabi.lib().brb_fp32_peak_flops(local_rank, C.byref(fl), C.byref(ms)), "brb_fp32_peak_flops")
    fp32_peak = fl.value

    policy = policy_params = None
    if args.actions == "policy":
        from balance_robot_b200.ppo import MlpPolicy
        torch.manual_seed(0)
        policy = MlpPolicy().to(dev)
        policy_params = policy.pack_params()
    obs_t = env.reset()

    def one_step(k):
        nonlocal obs_t
        if policy is None:
            obs_t = env.step(acts[k % nact])[0]
        else:
            _, _, _, a = policy.act_fused(obs_t, generator=gen, params=policy_params)      # one launch (csrc/brb_policy.cu)
            obs_t = env.step(a)[0]

    for k in range(args.spinup + args.warmup):
        one_step(k)
    torch.cuda.synchronize(dev)
    stats0 = env.stats()
    launches0 = env.num_launches()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)            # evict L2 (126 MB) between timed steps; not inside the event pair
        ev[k][0].record()
        one_step(k)
        ev[k][1].record()
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if sampler else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    launches = env.num_launches() - launches0
    stats1 = env.stats()
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = n * world * args.steps / (total_ms * 1e-3)

    # ---- e2e: the SB3 contract through the C-ABI host call (numpy in / numpy out, pinned staging, copies timed)
    e2e = None
    if not args.no_e2e:
        import numpy as np
        env_h = make_vec(args.env, n, device=dev, seed=args.seed + 1, env_id_offset=rank * n, output="numpy")
        env_h.reset()
        rng = np.random.default_rng(99 + rank)
        acts_h = [rng.uniform(-1, 1, (n, 2)).astype(np.float32) for _ in range(4)]
        e2e_steps = max(10, min(args.steps, 50))
        for k in range(args.spinup + 3):
            env_h.step(acts_h[k % 4])
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        n_done_rows = 0
        for k in range(e2e_steps):
            obs_h, rew_h, done_h, infos_h = env_h.step(acts_h[k % 4])
            n_done_rows += len(infos_h._idx)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": n * world * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": n * 2 * 4,
               "d2h_bytes_per_step": n * (6 * 4 + 4 + 1) + 4 + 40 * n_done_rows // e2e_steps, "steps": e2e_steps,
               "path": "BalanceVecEnv(output='numpy').step -> brb_env_step_host_compact (numpy actions in; numpy obs/reward/done out in "
                       "full, finished-episode records (terminal_observation, TimeLimit.truncated, Monitor r/l) compacted on the "
                       "device, 40 B per finished env; SB3 infos list built lazily)"}
        env_h.close()

    # ---- sustained: >= --sustained-s seconds of back-to-back steps in the same run (same per-step events, L2 flushed between)
    sustained = None
    if args.sustained_s > 0:
        ksus = max(args.steps, int(args.sustained_s / (total_ms / args.steps * 1e-3)) + 1)
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ksus)]
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        sampler2 = ClockSampler(local_rank) if rank == 0 else None
        w0 = time.perf_counter()
        for k in range(ksus):
            flush.fill_(k & 0xFF)
            ev2[k][0].record()
            one_step(k)
            ev2[k][1].record()
        torch.cuda.synchronize(dev)
        w1 = time.perf_counter() - w0
        clocks2 = sampler2.stop() if sampler2 else None
        sus_ms = sum(a.elapsed_time(b) for a, b in ev2)
        if dist is not None:
            t = torch.tensor([sus_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sus_ms = float(t.item())
        sustained = {"value": n * world * ksus / (sus_ms * 1e-3), "unit": UNIT, "steps": ksus, "ms_per_step": sus_ms / ksus,
                     "device_s": sus_ms * 1e-3, "wall_s": w1, "clocks": clocks2}

    # ---- policy rollout (BASELINE configs[2] style): on-device MlpPolicy inference inside every timed step, same shard
    policy_rollout = None
    if args.actions == "random" and args.ppo_envs_per_gpu > 0:
        from balance_robot_b200.ppo import MlpPolicy
        torch.manual_seed(0)
        pol = MlpPolicy().to(dev)
        pp = pol.pack_params()
        o = obs_t
        for k in range(3):
            o = env.step(pol.act_fused(o, generator=gen, params=pp)[3])[0]
        kp = max(10, args.steps)
        evp = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(kp)]
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        for k in range(kp):
            flush.fill_(k & 0xFF)
            evp[k][0].record()
            o = env.step(pol.act_fused(o, generator=gen, params=pp)[3])[0]
            evp[k][1].record()
        torch.cuda.synchronize(dev)
        pms = sum(a.elapsed_time(b) for a, b in evp)
        if dist is not None:
            t = torch.tensor([pms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pms = float(t.item())
        policy_rollout = {"value": n * world * kp / (pms * 1e-3), "unit": UNIT, "steps": kp, "ms_per_step": pms / kp,
                          "policy": "MlpPolicy 6-64-64-2 tanh (random init, seed 0), sampled on the device every step (brb_policy_act: 1 launch)"}
    env.close()

    # ---- ppo: rollout + update with the gradient all-reduce at N ranks (BASELINE configs[4]: Env01-v2, 1M envs per GPU)
    ppo_rec = ppo_sub_record(args, rank, world, dev, dist) if args.ppo_envs_per_gpu > 0 else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    sub = stats1["substeps"] - stats0["substeps"]
    csub = stats1["contact_substeps"] - stats0["contact_substeps"]
    active = csub / max(1, sub)
    kernel_ms = total_ms / args.steps
    # roofline of the dominant (only) kernel, per launch = one shard step.  `achieved` / `frac` use SURVEY.md 8(d)'s preferred
    # contact-weighted count F = 250 (1000 + 2700 active_fraction) FLOP per robot-step with the contact-active fraction measured by
    # the kernel's own counters in this run; `executed` is what the hardware did: (2 FFMA + FMUL + FADD thread instructions of the
    # committed ncu capture of this command) per robot-step x robots / this run's kernel time.
    occ_flop = FRAME_SKIP * (FLOP_PER_SUBSTEP_AIR + (FLOP_PER_SUBSTEP_CONTACT - FLOP_PER_SUBSTEP_AIR) * active) * n
    achieved = occ_flop / (kernel_ms * 1e-3) / 1e12
    all_contact = ALG_FLOP_PER_ENV_STEP * n / (kernel_ms * 1e-3) / 1e12
    try:
        peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
    except Exception:
        peaks = {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    ncu = load_ncu_summary()
    executed = None
    traffic = None
    if ncu:
        per_robot = (2 * ncu["thread_inst_ffma"] + ncu["thread_inst_fmul"] + ncu["thread_inst_fadd"]) / ncu["robots"]
        executed = {"flop_per_env_step": per_robot, "achieved": per_robot * n / (kernel_ms * 1e-3) / 1e12,
                    "frac": per_robot * n / (kernel_ms * 1e-3) / fp32_peak, "pipe_fma_pct": ncu.get("pipe_fma_pct"),
                    "issue_active_pct": ncu.get("issue_active_pct"), "lanes_active": ncu.get("lanes_active"),
                    "registers": ncu.get("registers"), "source": ncu["source"],
                    "note": "counters from the committed single-launch ncu capture (same command, 65,536 robots), time from this run"}
        traffic = ncu["dram_bytes"] / ncu["robots"] * n
    roofline = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak / 1e12, "unit": "TFLOP/s",
                "frac": achieved / (fp32_peak / 1e12),
                "formula": "250 x (1000 + 2700 x contact_active_fraction) FLOP per robot-step (SURVEY.md 8d), contact_active_fraction measured in this run",
                "traffic": traffic,
                "traffic_source": (f"dram__bytes_read.sum + dram__bytes_write.sum per launch from {ncu['source']}, scaled to this shard; "
                                   "algorithmic bytes = 290 B per robot-step") if ncu else None,
                "peak_source": "FFMA probe kernel measured in this run (brb_fp32_peak_flops); MEASURED_PEAKS.json has no FP32 entry; "
                               f"nominal {NOMINAL_FP32_TFLOPS:.1f}",
                "contact_active_fraction": active,
                "executed": executed,
                "all_contact_count": {"flop_per_env_step": ALG_FLOP_PER_ENV_STEP, "achieved": all_contact, "frac": all_contact / (fp32_peak / 1e12),
                                      "note": "auxiliary: charges 4 contacts to every substep, also the airborne ones; overstates the work"},
                "hbm": {"achieved_gbs": ALG_BYTES_PER_ENV_STEP * n / (kernel_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
                "kernel": "brb_step_kernel<Env01_v2>", "kernel_ms": kernel_ms}
    cpu_baseline = None
    if not args.no_cpu_baseline and world >= 1:
        threads = args.cpu_threads or (os.cpu_count() or 1)
        v, dt, ncpu = cpu_leg(args, 120, 10, threads)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{ncpu} envs (64 per thread) x 120 steps after {args.spinup} spin-up + 10 warm-up ({dt:.1f} s timed), fp64 oracle, replayed Philox draws"}
    line = {"metric": METRIC.replace(ENV_ID, args.env), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, world), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "wall_s": wall, "sustained": sustained,
            "policy_rollout": policy_rollout, "ppo": ppo_rec,
            "solver": {"nonconverged_substeps": stats1["nonconverged"] - stats0["nonconverged"],
                       "unsupported_pose_steps": stats1["unsupported"] - stats0["unsupported"],
                       "solves_per_contact_substep": (stats1["solves"] - stats0["solves"]) / max(1, csub),
                       "episodes": stats1["episodes"] - stats0["episodes"]}}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def ppo_sub_record(args, rank, world, dev, dist):
    """Trained env-steps/s of on-device PPO (rollout with policy inference + update) at `world` ranks: envs sharded by rank, one
    flat-gradient all-reduce (NCCL) per minibatch.  Device-timed, max over ranks; the update is also split into its three pieces
    (gradient kernels / all-reduce / clip + Adam kernel), each timed alone over the same minibatch shapes."""
    import ctypes as C
    import torch
    from balance_robot_b200 import make_vec, _cabi
    from balance_robot_b200.ppo import PPO, PPOConfig
    n = args.ppo_envs_per_gpu
    cfg = PPOConfig(n_steps=16, n_epochs=10, n_minibatches=4, seed=0)
    env = make_vec(args.env, n, device=dev, seed=args.seed + 7, env_id_offset=rank * n)
    agent = PPO(env, cfg, device=dev, rank=rank, world_size=world)
    agent.collect_rollouts(); agent.train()                   # warm-up iteration (also past the all-airborne start)

    def maxr(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    roll_ms = upd_ms = 0.0
    for _ in range(args.ppo_iters):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e[0].record(); agent.collect_rollouts(); e[1].record(); agent.train(); e[2].record()
        torch.cuda.synchronize(dev)
        roll_ms += e[0].elapsed_time(e[1]); upd_ms += e[1].elapsed_time(e[2])
    roll_ms, upd_ms = maxr(roll_ms / args.ppo_iters), maxr(upd_ms / args.ppo_iters)
    # pieces of one minibatch update, each alone
    total = cfg.n_steps * n
    mb = total // cfg.n_minibatches
    L, stream = _cabi.lib(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    b = agent.buf
    obs, act = b["obs"].reshape(total, 6), b["actions"].reshape(total, 2)
    oldlp, adv, ret = b["logp"].reshape(total), b["adv"].reshape(total).contiguous(), b["ret"].reshape(total).contiguous()
    idx = torch.randperm(total, device=dev)[:mb]
    astats = torch.tensor([0.0, 1.0], device=dev)
    g, st = torch.zeros_like(agent._gflat), torch.zeros(4, device=dev)
    p2, m2, v2 = agent._pflat.clone(), agent._m.clone(), agent._v.clone()

    def timed(fn, reps):
        fn(); torch.cuda.synchronize(dev)
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        a.record()
        for _ in range(reps):
            fn()
        z.record(); torch.cuda.synchronize(dev)
        return maxr(a.elapsed_time(z) / reps)
    grad_ms = timed(lambda: _cabi.check(L.brb_ppo_grad(p2.data_ptr(), obs.data_ptr(), act.data_ptr(), oldlp.data_ptr(), adv.data_ptr(), ret.data_ptr(),
                                                       idx.data_ptr(), mb, astats.data_ptr(), cfg.clip_range, cfg.vf_coef, cfg.ent_coef,
                                                       g.data_ptr(), st.data_ptr(), stream), "brb_ppo_grad"), 5)
    split = {"grad_kernels": grad_ms}
    if agent._comm is not None:
        # the fused peer-memory kernel (all-reduce over NVLink + clipping + Adam): every rank calls it in lockstep with the next step numbers
        def fused():
            agent._adam_t += 1
            _cabi.check(L.brb_comm_allreduce_adam(agent._comm, p2.data_ptr(), m2.data_ptr(), v2.data_ptr(), 3e-4, 0.9, 0.999, 1e-5, agent._adam_t,
                                                  0.5, None, stream), "brb_comm_allreduce_adam")
        split["allreduce_clip_adam_kernel"] = timed(fused, 20)
        split["path"] = "one kernel: P2P loads of every rank's gradient over NVLink, clipping, Adam (brb_comm_allreduce_adam)"
        assert L.brb_comm_fault(agent._comm) == 0
    else:
        split["all_reduce"] = timed(lambda: dist.all_reduce(g), 20) if dist is not None else 0.0
        split["clip_adam_kernel"] = timed(lambda: _cabi.check(L.brb_adam_clip_step(p2.data_ptr(), g.data_ptr(), m2.data_ptr(), v2.data_ptr(), p2.numel(),
                                                                                   3e-4, 0.9, 0.999, 1e-5, 1, 0.5, 1.0 / world, None, stream),
                                                              "brb_adam_clip_step"), 20)
        split["path"] = "NCCL all_reduce + brb_adam_clip_step" if dist is not None else "single rank: brb_adam_clip_step"
    nupd = cfg.n_epochs * cfg.n_minibatches
    out = {"value": n * world * cfg.n_steps / ((roll_ms + upd_ms) * 1e-3), "unit": "trained env-steps/s", "envs_per_gpu": n, "n_gpus": world,
           "rollout_ms": roll_ms, "update_ms": upd_ms, "rollout_env_steps_per_s": n * world * cfg.n_steps / (roll_ms * 1e-3),
           "update_samples_per_s": total * world * cfg.n_epochs / (upd_ms * 1e-3),
           "update_split_ms_per_minibatch": dict(split, measured_whole=upd_ms / nupd),
           "grad_kernel": "brb_ppo_grad_tc_kernel: tcgen05.mma kind::f16, bf16 hi/lo split x 3 passes, fp32 accumulators in TMEM",
           "config": f"{args.env}, {n} envs per GPU, n_steps {cfg.n_steps}, {cfg.n_epochs} epochs x {cfg.n_minibatches} minibatches of {mb} samples "
                     f"per rank, one {4 * agent._pflat.numel()} B gradient all-reduce per minibatch ({split["path"]})",
           "iterations_timed": args.ppo_iters}
    agent.close()
    env.close()
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000)] + sys.argv
        raise SystemExit(subprocess.call(cmd))
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
