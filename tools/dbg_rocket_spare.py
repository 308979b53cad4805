import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pyflyt_b200.gym_envs import RocketLandingVecEnv
for cr in (False, True):
    for rep in range(2):
        outs = []
        for inline in (False, True):
            env = RocketLandingVecEnv(num_envs=4096, seed=7, inline_reset=inline, max_duration_seconds=0.4, randomize_drop=True, contact_response=cr)
            env.reset()
            for k in range(70):
                env.rollout(1)
            torch.cuda.synchronize()
            outs.append((env.aviary.obs.clone(), env.aviary.reward.clone(), env.aviary.state_tensor.clone()))
            env.close()
        a, b = outs
        d = [(x.double() - y.double()).abs() for x, y in zip(a, b)]
        print(f"contact_response={cr} rep={rep}: obs max diff {d[0].max().item():.3e} (n={int((d[0]>0).sum())}), reward {d[1].max().item():.3e}, state {d[2].max().item():.3e} (rows {sorted(set((d[2]>0).nonzero()[:,0].tolist()))[:12]}, envs {int((d[2]>0).any(0).sum())})", flush=True)
