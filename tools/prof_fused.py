import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pyflyt_b200.gym_envs.quadx_hover_env import QuadXHoverVecEnv
env = QuadXHoverVecEnv(num_envs=65536, seed=0)
env.reset()
for _ in range(6):
    env.rollout(16)
torch.cuda.synchronize()
