import os, sys, subprocess, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
mode = sys.argv[1]
ENV = {"fused4": dict(PFB_FUSED_MIN="4"), "fused4b": dict(PFB_FUSED_MIN="4"), "fused1": dict(PFB_FUSED_MIN="1"), "fused1noreq": dict(PFB_FUSED_MIN="1", PFB_FUSED_NOREQ="1"),
       "step": dict(PFB_FUSED_MIN="1000"), "stepb": dict(PFB_FUSED_MIN="1000")}
if mode == "driver":
    for m in ENV:
        subprocess.run([sys.executable, __file__, m], check=True, env=dict(os.environ, **ENV[m]))
    D = {m: torch.load(f"/tmp/dbg_{m}.pt") for m in ENV}
    for x, y in (("fused4", "fused4b"), ("step", "stepb"), ("fused1noreq", "step"), ("fused1", "step"), ("fused4", "step"), ("fused4", "fused1")):
        d = [(a.double() - b.double()).abs() for a, b in zip(D[x], D[y])]
        print(x, "vs", y, "state envs differing per checkpoint:", [int((q.amax(dim=(1, 3)) > 0).sum()) for q in d])
else:
    from pyflyt_b200.gym_envs.quadx_hover_env import QuadXHoverVecEnv
    env = QuadXHoverVecEnv(num_envs=65536, seed=21)
    env.reset()
    out = []
    for w in range(10):
        if mode.startswith("fused4"):
            env.rollout(4)
        else:
            for _ in range(4):
                env.rollout(1)
        out.append(env.aviary.state_tensor.clone().cpu())
    torch.save(out, f"/tmp/dbg_{mode}.pt")
