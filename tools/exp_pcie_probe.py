"""How fast does THIS box move a step's results over PCIe?  DMA copies (cudaMemcpyAsync from / to pinned memory) at the sizes of
one QuadX-Hover step (1 MiB of actions in, 5.9 MB of observations / rewards / flags out) and at 64 MiB, next to the zero-copy
number `bench.py` prints.  Experiment harness (profiles/), not a product path."""
import json
import time

import torch

dev = torch.device("cuda", 0)
out = {}
for name, nbytes in (("actions_1MiB", 1 << 20), ("results_5.9MB", 65536 * 90), ("chunk_688KB", 8192 * 84), ("big_64MiB", 64 << 20)):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for direction in ("d2h", "h2d"):
        best = 1e9
        for rep in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            if direction == "d2h":
                h.copy_(d, non_blocking=True)
            else:
                d.copy_(h, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[f"{name}_{direction}"] = {"us": best * 1e3, "GBps": nbytes / (best * 1e-3) / 1e9}
# 8 chunked D2H copies back to back on one stream (the shape of a chunk-pipelined host entry)
nb = 8192 * 84
h = torch.empty(8 * nb, dtype=torch.uint8).pin_memory()
d = torch.empty(8 * nb, dtype=torch.uint8, device=dev)
best = 1e9
for rep in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for c in range(8):
        h[c * nb:(c + 1) * nb].copy_(d[c * nb:(c + 1) * nb], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
out["8_chunks_688KB_d2h"] = {"us": best * 1e3, "GBps": 8 * nb / (best * 1e-3) / 1e9}
print(json.dumps(out))
