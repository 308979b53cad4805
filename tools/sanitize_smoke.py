"""Small end-to-end pass over every env kind for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pyflyt_b200.gym_envs import FixedwingWaypointsVecEnv, QuadXHoverVecEnv, QuadXWaypointsVecEnv, RocketLandingVecEnv
from pyflyt_b200.pz_envs import MAFixedwingDogfightSplitEnv, MAFixedwingDogfightVecEnv, MAQuadXHoverVecEnv

def drive(env, n):
    env.reset()
    for _ in range(n):
        env.rollout(1)
    torch.cuda.synchronize()
    env.close()

drive(QuadXHoverVecEnv(num_envs=1000, seed=1, max_duration_seconds=0.2), 30)          # ragged last CTA, many autoresets
drive(QuadXHoverVecEnv(num_envs=96, seed=1, flight_mode=6, angle_representation="euler"), 10)
# fused rollout (k_hover_rollout + spare top-up + the hand-over both ways) and 3-step episodes (every lane of a warp resets at once:
# the lanes beyond the staging slots take the direct path)
env = QuadXHoverVecEnv(num_envs=1000, seed=2, max_duration_seconds=0.2)
env.reset()
for n in (1, 16, 1, 5, 16):
    env.rollout(n)
torch.cuda.synchronize()
env.close()
drive(QuadXHoverVecEnv(num_envs=200, seed=3, max_duration_seconds=0.05), 12)
drive(QuadXWaypointsVecEnv(num_envs=500, seed=1, use_yaw_targets=True, max_duration_seconds=0.3), 20)
drive(FixedwingWaypointsVecEnv(num_envs=300, seed=1, max_duration_seconds=0.3), 20)
drive(RocketLandingVecEnv(num_envs=300, seed=1, max_duration_seconds=0.3), 20)
drive(MAFixedwingDogfightVecEnv(num_arenas=100, team_size=2, seed=1, max_duration_seconds=0.3), 20)
ma = MAQuadXHoverVecEnv(num_arenas=70, seed=1, flight_dome_size=2.0)
ma.reset()
for _ in range(40):
    ma.step(torch.rand(ma.num_agents, 4, device=ma.device) * 2 - 1)
torch.cuda.synchronize()
ma.close()
for ex in ("nccl", "peer", "peer-signal"):
    env = MAFixedwingDogfightSplitEnv(64, seed=1, exchange=ex)
    env.reset()
    for _ in range(5):
        env.step(torch.zeros(env.n_local, 4, device=env.device))
    torch.cuda.synchronize()
    env.close()
print("SANITIZE_SMOKE_DONE")
