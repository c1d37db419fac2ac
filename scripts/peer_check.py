"""Check of the fused reduce + cross-rank sum of the streaming pass (pla_stream_pass_peer_f64) against the sum of the
per-rank results, bit for bit.  Run under torchrun; `--same-gpu` puts every rank on cuda:0 (the ranks then exchange
through CUDA IPC mappings of the same device and time-slice it), so the path is testable on a one-GPU box:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/peer_check.py --same-gpu
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parla_b200 import kernels as K                   # noqa: E402
from parla_b200.parallel import PeerComm              # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", 0 if "--same-gpu" in sys.argv else local)
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")
    comm = PeerComm(dist.group.WORLD, dev, 8193)
    assert comm.ok, "CUDA IPC mapping of the exchange buffers failed"
    ok = True
    shapes = [(3000, 64), (5000, 500), (2048, 2048), (4096, 1000), (1000, 8192), (3000, 64), (3000, 64), (777, 333)]
    stop = torch.ones(1, dtype=torch.int32, device=dev)
    for trial, (m, n) in enumerate(shapes):
        g = torch.Generator(device=dev).manual_seed(100 * trial + rank)
        A = torch.randn(m + 16 * rank, n, dtype=torch.float64, device=dev, generator=g)     # ragged shards
        w = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
        u0 = torch.randn(A.shape[0], dtype=torch.float64, device=dev, generator=g)
        for flags in (K.PASS_DOT | K.PASS_AXPY, K.PASS_AXPY, K.PASS_DOT):
            ul, uf = u0.clone(), u0.clone()
            kw = dict(w=w if flags & K.PASS_DOT else None, sa=0.5, su=-1.0, flags=flags)
            z_loc = K.stream_pass(A, u=ul, **kw).cpu()
            z_fused = K.stream_pass(A, u=uf, comm=comm, **kw).cpu()
            parts = [torch.empty_like(z_loc) for _ in range(world)]
            dist.all_gather(parts, z_loc)
            ref = np.zeros(n + 1)
            for p in parts:                                   # rank order, starting from 0.0: the kernel's order
                ref = ref + p.numpy()
            same = np.array_equal(ref, z_fused.numpy()) and torch.equal(ul, uf)
            ok = ok and same
            if rank == 0:
                print(f"{'OK  ' if same else 'FAIL'} fused pass {m}x{n} flags={flags}: max |diff| = "
                      f"{np.max(np.abs(ref - z_fused.numpy())):.3e}", flush=True)
        if trial == 2:                                        # a call every rank skips (LSQR has stopped): epoch moves on
            before = torch.full((n + 1,), 7.0, dtype=torch.float64, device=dev)
            K.stream_pass(A, w=w, u=u0.clone(), zss=before, flags=K.PASS_DOT | K.PASS_AXPY, istop=stop, comm=comm)
            same = bool((before == 7.0).all())
            ok = ok and same
            if rank == 0:
                print(f"{'OK  ' if same else 'FAIL'} stopped call leaves zss untouched", flush=True)
    flag = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("PEER_CHECK", "PASS" if float(flag) == 1.0 else "FAIL", f"world={world} device={dev}", flush=True)
    torch.cuda.synchronize()
    comm.close()
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
