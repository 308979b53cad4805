import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pyflyt_b200.gym_envs.quadx_hover_env import QuadXHoverVecEnv
n = 65536
a = QuadXHoverVecEnv(num_envs=n, seed=21); b = QuadXHoverVecEnv(num_envs=n, seed=21)
a.reset(); b.reset()
seen = set()
for w in range(24):
    prev = b.aviary.state_tensor.clone()
    a.rollout(4)
    hist = []
    for _ in range(4):
        b.rollout(1)
        hist.append((b.aviary.state_tensor.clone(), b.aviary.obs.clone(), b.aviary.term.clone(), b.aviary.trunc.clone()))
    sa, sb = a.aviary.state_tensor, b.aviary.state_tensor
    d = (sa.double() - sb.double()).abs()
    envs = (d.amax(dim=(1, 3)) > 0).nonzero().tolist()
    new = [(t, l) for t, l in envs if (t, l) not in seen]
    if new:
        print(f"window {w} (steps {4*w}..{4*w+3}): {len(new)} new envs differ")
        for t, l in new[:3]:
            seen.add((t, l))
            i = t * 32 + l
            rows = (d[t, :, l, :] > 0).nonzero().tolist()
            print("  env", i, "rows", [4 * g + c for g, c in rows])
            print("   before window (b):", [round(v, 9) for v in prev[t, :, l, :].reshape(-1)[:36].tolist()])
            for k, (st, ob, te, tr) in enumerate(hist):
                print(f"   after b step {k}: z={st[t,0,l,2].item():.6f} w=({st[t,2,l,2].item():.4f},{st[t,2,l,3].item():.4f},{st[t,3,l,0].item():.4f}) thr={st[t,3,l,1].item():.4f} term={int(te[i])} trunc={int(tr[i])} flags={st[t,4,l,2].view(torch.int32).item()} step={st[t,4,l,1].view(torch.int32).item()}")
            print("   fused end :", [round(v, 12) for v in sa[t, :, l, :].reshape(-1)[24:36].tolist()])
            print("   step  end :", [round(v, 12) for v in sb[t, :, l, :].reshape(-1)[24:36].tolist()])
        for t, l in new[3:]:
            seen.add((t, l))
