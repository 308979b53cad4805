"""Step-kernel timings for the other BASELINE.json configs (parity-test cases, not bench.py lines): Fixedwing-Waypoints,
Rocket-Landing, MAFixedwingDogfight (fused arena kernel, and the split all-gather variant).  Same method as bench.py:
L2 flushed between steps, per-step CUDA-event pairs.  One JSON line per workload.

    python tools/bench_workloads.py [--steps 100] [--warmup 5]
    python -m torch.distributed.run --nproc-per-node 2 ... tools/bench_workloads.py --only dogfight-split
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def time_steps(step, K, W, dev, world):
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    for _ in range(W):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for k in range(K):
        flush.fill_(float(k))
        ev[k][0].record()
        step()
        ev[k][1].record()
    torch.cuda.synchronize(dev)
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer", "peer-signal"], help="dogfight-split: NCCL all-gather or peer stores + barrier")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from pyflyt_b200.gym_envs.fixedwing_waypoints_env import FixedwingWaypointsVecEnv
    from pyflyt_b200.gym_envs.rocket_landing_env import RocketLandingVecEnv
    from pyflyt_b200.pz_envs import MAFixedwingDogfightSplitEnv, MAFixedwingDogfightVecEnv

    K, W = args.steps, args.warmup
    out = []

    def report(name, units, ms, launches, extra=None):
        if rank == 0:
            line = {"workload": name, "metric": "env-steps/s", "value": world * units * K / (ms * 1e-3), "ms_per_step": ms / K, "n_gpus": world,
                    "units_per_gpu": units, "steps": K, "gpu_launches_per_step": launches, "l2": "flushed between steps"}
            line.update(extra or {})
            print(json.dumps(line), flush=True)

    want = lambda n: not args.only or args.only == n  # noqa: E731
    if want("fixedwing-waypoints"):
        n = 16384
        env = FixedwingWaypointsVecEnv(num_envs=n, seed=1, device=dev, env_offset=rank * n)
        env.reset()
        report("Fixedwing-Waypoints-v4, 16384 envs/GPU (configs[2]), random actions, NEXT_STEP autoreset", n, time_steps(lambda: env.rollout(1), K, W, dev, world), 1)
        env.close()
    if want("quadx-waypoints"):
        from pyflyt_b200.gym_envs import QuadXWaypointsVecEnv

        n = 65536
        env = QuadXWaypointsVecEnv(num_envs=n, seed=1, device=dev, env_offset=rank * n)
        env.reset()
        report("QuadX-Waypoints-v4, 65536 envs/GPU (SURVEY 8f #1), mode 0, random actions, NEXT_STEP autoreset (inline warm-ups)", n,
               time_steps(lambda: env.rollout(1), K, W, dev, world), 1)
        env.close()
    if want("rocket-landing"):
        n = 16384
        env = RocketLandingVecEnv(num_envs=n, seed=1, device=dev, env_offset=rank * n)
        env.reset()
        report("Rocket-Landing-v4, 16384 envs/GPU (configs[3]), random actions, NEXT_STEP autoreset", n, time_steps(lambda: env.rollout(1), K, W, dev, world), 1)
        env.close()
    if want("dogfight"):
        arenas = 8192
        env = MAFixedwingDogfightVecEnv(num_arenas=arenas, seed=1, device=dev, env_offset=rank * arenas * 2)
        env.reset()
        report("MAFixedwingDogfight, 8192 arenas x 2 agents per GPU (configs[4]), arena-sharded fused kernel (no collective); value counts agent-steps",
               2 * arenas, time_steps(lambda: env.rollout(1), K, W, dev, world), 1)
        env.close()
    if want("dogfight-split"):
        arenas = 8192 * world
        env = MAFixedwingDogfightSplitEnv(arenas, seed=1, device=dev, exchange=args.exchange)
        env.reset()
        act = torch.rand(env.n_local, 4, device=dev) * 2 - 1
        c0 = env.collectives
        ms = time_steps(lambda: env.step(act), K, W, dev, world)
        report(f"MAFixedwingDogfight split: {arenas} arenas x 2 agents over {world} rank(s), exchange={args.exchange} every Aviary step; value counts agent-steps",
               env.n_local, ms, 8, {"collectives_per_step": 4, "payload_bytes_per_rank_per_collective": env.n_local * 80})
        env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
