"""`sb_rl.py`-compatible command line for the B200 path (reference: src/sb_rl.py:492-593).

    python sb_rl.py -a PPO train -e Env01-v2                       # same flags as the reference (README.md:58)
    python sb_rl.py -a PPO -m models/Env01-v2_PPO/best_model.zip train -e Env01-v3   # fine-tune (README.md:62)
    torchrun --nproc-per-node 8 sb_rl.py -a PPO train -e Env01-v2 --num-envs 1048576   # envs sharded over 8 GPUs

Kept from the reference: the click group with -a/--algorithm (required) and -m/--model, `train -e`, the folders
models/ logs/ movies/, model naming `models/<env>_<algo>/best_model.zip` and `<env>_<algo>_cp__<steps>_steps.zip`,
checkpoint / eval cadence (40,000 / 20,000 steps, scaled up to whole rollouts), the 1e10-step default horizon,
the reward-threshold stop at 6000.  New: --num-envs / --total-timesteps / --device / --seed.
Only PPO is built (the other SB3 algorithms the reference accepts are replay-buffer methods outside the north star);
the viewer / ONNX / TFLite commands are out of scope and say so.
"""
from __future__ import annotations

import logging
import os
import pathlib

import click
import torch

from . import make_vec, registry
from .ppo import PPO, PPOConfig, evaluate_policy

logging.basicConfig(format="%(levelname)s:%(message)s", level=logging.INFO)
MODEL_DIR, LOG_DIR, MOVIE_DIR = "models", "logs", "movies"        # sb_rl.py:35-37
SUPPORTED = ("PPO",)
SB3_ALGOS = ("PPO", "DDPG", "SAC", "TD3", "A2C")


def _make_folders():
    for d in (MODEL_DIR, LOG_DIR, MOVIE_DIR):
        os.makedirs(d, exist_ok=True)


@click.group()
@click.option("-a", "--algorithm", required=True, type=str, help="Name of Stable Baselines3 algorithm (eg; PPO)")
@click.option("-m", "--model", "model_file", required=False, type=click.Path(), help="Existing model to load")
@click.pass_context
def cli(ctx, algorithm, model_file):
    if algorithm not in SB3_ALGOS:
        raise click.BadParameter(f"{algorithm} is not a Stable Baselines3 algorithm")      # sb_rl.py:575-581
    if algorithm not in SUPPORTED:
        raise click.UsageError(f"{algorithm}: only PPO runs on the B200 path (north star: PPO against the batched envs)")
    ctx.ensure_object(dict)
    ctx.obj["algorithm"], ctx.obj["model_file"] = algorithm, model_file


@cli.command(help="Train a model on the batched B200 environments")
@click.option("-e", "--environment", required=True, type=str, help="ID of Environment to train against")
@click.option("--num-envs", default=4096, show_default=True, type=int, help="robots per GPU")
@click.option("--total-timesteps", default=int(1e10), show_default=True, type=int)
@click.option("--n-steps", default=32, show_default=True, type=int, help="rollout length per env")
@click.option("--device", default=None, type=str)
@click.option("--seed", default=0, type=int)
@click.pass_context
def train(ctx, environment, num_envs, total_timesteps, n_steps, device, seed):
    algo, model_file = ctx.obj["algorithm"], ctx.obj["model_file"]
    spec = registry.spec(environment)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = device or f"cuda:{local}"
    env = make_vec(environment, num_envs, device=device, seed=seed, env_id_offset=rank * num_envs)
    cfg = PPOConfig(n_steps=n_steps, seed=seed)
    if model_file:
        if not pathlib.Path(model_file).exists():
            raise RuntimeError(f"model file {model_file} does not exist")                     # sb_rl.py:527
        agent = PPO.load(model_file, env, cfg, device=device, rank=rank, world_size=world)
        agent.num_timesteps = 0
    else:
        agent = PPO(env, cfg, device=device, rank=rank, world_size=world)

    run_dir = pathlib.Path(MODEL_DIR) / f"{environment}_{algo}"
    writer = None
    if rank == 0:
        _make_folders()
        try:
            from torch.utils.tensorboard import SummaryWriter
            k = 1
            while (pathlib.Path(LOG_DIR) / f"{environment}_{algo}_{k}").exists():
                k += 1
            writer = SummaryWriter(str(pathlib.Path(LOG_DIR) / f"{environment}_{algo}_{k}"))     # tb_log_name, sb_rl.py:554
        except Exception as exc:  # tensorboard missing: keep training, say so
            logging.warning("tensorboard unavailable (%s); scalars go to stdout only", exc)
    eval_env = make_vec(environment, 5, device=device, seed=seed + 7919) if rank == 0 else None     # n_eval_episodes=5, sb_rl.py:540
    per_iter = n_steps * num_envs * world
    state = {"best": -float("inf"), "next_eval": 20000, "next_ckpt": 40000, "no_improve": 0, "evals": 0}

    def callback(agent: PPO, rec) -> bool:
        stop = False
        if rank == 0:
            if agent.num_timesteps >= state["next_ckpt"]:                                           # CheckpointCallback, sb_rl.py:545-550
                agent.save(run_dir / f"{environment}_{algo}_cp__{agent.num_timesteps}_steps.zip")
                state["next_ckpt"] = agent.num_timesteps + max(40000, per_iter)
            if agent.num_timesteps >= state["next_eval"]:                                           # EvalCallback, sb_rl.py:536-543
                mean_r, std_r, lens = evaluate_policy(agent.policy, eval_env, 5, True, spec.max_episode_steps)
                state["evals"] += 1
                logging.info("Eval num_timesteps=%d, episode_reward=%.2f +/- %.2f", agent.num_timesteps, mean_r, std_r)
                if writer is not None:
                    writer.add_scalar("eval/mean_reward", mean_r, agent.num_timesteps)
                if mean_r > state["best"]:
                    state["best"], state["no_improve"] = mean_r, 0
                    agent.save(run_dir / "best_model.zip")
                    if mean_r >= spec.reward_threshold:                                             # StopTrainingOnRewardThreshold(6000), sb_rl.py:529
                        stop = True
                else:
                    state["no_improve"] += 1
                    if state["evals"] > 10000 and state["no_improve"] > 5:                          # StopTrainingOnNoModelImprovement, sb_rl.py:530-534
                        stop = True
                state["next_eval"] = agent.num_timesteps + max(20000, per_iter)
        if world > 1:
            flag = torch.tensor([1.0 if stop else 0.0], device=device)
            agent._all_reduce_(flag)
            stop = bool(flag.item() > 0)
        return not stop

    agent.learn(total_timesteps, callback=callback, writer=writer)
    if rank == 0:
        agent.save(run_dir / f"{environment}_{algo}_final.zip")
        if writer is not None:
            writer.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def _out_of_scope(name, why):
    @cli.command(name=name, help=f"(not built: {why})")
    @click.option("-e", "--environment", required=False, type=str)
    def _cmd(environment):
        raise click.UsageError(f"`{name}` is outside the B200 hot-path scope: {why}")
    return _cmd


@cli.command(help="Test the current model (headless: the reference opens the MuJoCo viewer, sb_rl.py:136-182)")
@click.option("-e", "--environment", required=True, type=str, help="id of the environment (eg; Env01-v1)")
@click.option("--show-io", is_flag=True, default=False, help="log model inputs and outputs")
@click.option("--show-i", is_flag=True, default=False, help="log model inputs to std out in Python array syntax")
@click.option("--episodes", default=20, show_default=True, type=int, help="episodes to run (the reference loops until interrupted)")
@click.option("--device", default="cuda:0", type=str)
@click.pass_context
def test(ctx, environment, show_io, show_i, episodes, device):
    algo, model_file = ctx.obj["algorithm"], ctx.obj["model_file"]
    spec = registry.spec(environment)
    if model_file is None:                                                                          # default name, sb_rl.py:147-149
        model_file = os.path.join(MODEL_DIR, f"{environment}_{algo}", "best_model.zip")
    if not os.path.isfile(model_file):
        raise RuntimeError(f"Could not open model file: {model_file}")                              # sb_rl.py:151-152
    logging.info("Starting test simulation")
    logging.info("Algorithm: %s", algo)
    logging.info("Environment: %s", environment)
    logging.info("Model: %s", model_file)
    n = max(1, min(episodes, 64))
    env = make_vec(environment, n, device=device, seed=12345)
    agent = PPO.load(model_file, env, PPOConfig(), device=device)
    obs = env.reset()
    returns, lengths, k = [], [], 0
    while len(returns) < episodes and k < 4 * spec.max_episode_steps:
        action, _ = agent.policy.predict(obs)                                                       # model.predict(obs), sb_rl.py:166
        if (show_io or show_i) and k % 30 == 0:
            row = [float(v) for v in obs[0].tolist()] + ([float(v) for v in action[0].tolist()] if show_io else [])
            logging.info(str(row) + ("," if show_i and not show_io else ""))
        obs, _, done, infos = env.step(action)
        for i in torch.nonzero(done).flatten().tolist():
            returns.append(float(infos.episode_return[i])); lengths.append(int(infos.episode_length[i]))
        k += 1
    returns, lengths = returns[:episodes], lengths[:episodes]
    if returns:
        logging.info("episodes %d  mean return %.2f  mean length %.1f  (min %d, max %d)", len(returns), sum(returns) / len(returns),
                     sum(lengths) / len(lengths), min(lengths), max(lengths))
        click.echo(f"episodes={len(returns)} mean_return={sum(returns) / len(returns):.3f} mean_length={sum(lengths) / len(lengths):.1f}")
    env.close()


_out_of_scope("convert", "ONNX export (sb_rl.py:86-133)")
_out_of_scope("test-onnx", "onnxruntime roll-out (sb_rl.py:185-247)")
_out_of_scope("test-tflite", "TFLite roll-out (sb_rl.py:250-306)")
_out_of_scope("test-tflite-quant", "quantised TFLite roll-out (sb_rl.py:309-364)")
_out_of_scope("test-tflite-arduino", "serial link to the Teensy (sb_rl.py:367-489)")


def main():
    cli(obj={})


if __name__ == "__main__":
    main()
