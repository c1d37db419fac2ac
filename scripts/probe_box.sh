#!/bin/bash
# What the GPU box looks like (host cores / RAM / GPU / PCIe), for DESIGN.md and bench sizing.
echo "nproc=$(nproc)"; free -g | head -2
nvidia-smi --query-gpu=index,name,memory.total,clocks.max.sm,clocks.max.mem,power.limit,pcie.link.gen.max,pcie.link.width.max --format=csv
python - <<'PY'
import os, numpy as np, scipy
print("affinity cores", len(os.sched_getaffinity(0)), "numpy", np.__version__, "scipy", scipy.__version__)
try:
    from threadpoolctl import threadpool_info
    for i in threadpool_info(): print(i.get("internal_api"), i.get("version"), i.get("num_threads"))
except Exception as e: print("threadpoolctl:", e)
import torch, time
x = torch.empty(1<<30, dtype=torch.uint8).pin_memory()
d = torch.empty(1<<30, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t=time.time(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); print("H2D pinned GB/s", 1.073741824/(time.time()-t))
t=time.time(); x.copy_(d, non_blocking=True); torch.cuda.synchronize(); print("D2H pinned GB/s", 1.073741824/(time.time()-t))
PY
