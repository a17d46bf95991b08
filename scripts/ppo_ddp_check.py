"""Multi-GPU check (torchrun): envs sharded by rank, fused PPO update, one gradient all-reduce per minibatch — through the fused
peer-memory all-reduce + clip + Adam kernel (default) or NCCL + brb_adam_clip_step (BRB_PPO_NO_P2P=1); the policy replicas must
stay bit-identical and the rollout statistics are summed over ranks.  Prints the update time of the path in use."""
import os, sys, time
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from balance_robot_b200 import make_vec
from balance_robot_b200.ppo import PPO, PPOConfig
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = make_vec("Env01-v2", n, device=f"cuda:{local}", seed=0, env_id_offset=rank * n)
agent = PPO(env, PPOConfig(n_steps=16, seed=3), device=f"cuda:{local}", rank=rank, world_size=world)
path = "peer-memory allreduce+adam kernel" if agent._comm is not None else "NCCL all_reduce + adam kernel"
for it in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    roll = agent.collect_rollouts(); torch.cuda.synchronize(); t1 = time.perf_counter()
    upd = agent.train(); torch.cuda.synchronize(); dist.barrier(); t2 = time.perf_counter()
    flat = agent.policy.pack_params()
    ref = flat.clone(); dist.broadcast(ref, 0)
    same = torch.tensor([float(torch.equal(flat, ref))], device=flat.device); dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"it {it}: world {world} x {n} envs, rollout {t1 - t0:.3f} s, update {t2 - t1:.3f} s, {world * n * 16 / (t2 - t0):.3e} env-steps/s trained, "
              f"episodes {roll['episodes']:.0f}, replicas identical: {bool(same.item())}, kl {upd['approx_kl']:.5f}, path: {path}, "
              f"param checksum {float(flat.double().abs().sum()):.9f}", flush=True)
    assert same.item() == 1.0
if agent._comm is not None:
    from balance_robot_b200 import _cabi
    assert _cabi.lib().brb_comm_fault(agent._comm) == 0, "a wait for a peer timed out"
agent.close()
dist.destroy_process_group()
