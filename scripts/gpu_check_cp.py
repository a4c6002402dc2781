#!/usr/bin/env python
"""In-kernel context-parallel attention (parallel/cp.py CPT5Encoder) vs the single-GPU T5 encoder on the full sequence.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 scripts/gpu_check_cp.py
"""
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from infomesh_b200.models.t5 import T5_SMALL, T5Model  # noqa: E402
from infomesh_b200.parallel import cp as CP  # noqa: E402
from infomesh_b200.parallel import dist as D  # noqa: E402

ctx = D.init()
rank, world, dev = ctx.rank, ctx.world, ctx.device
model = T5Model(T5_SMALL, device=dev, seed=11)
allok = True
for (B, S) in [(2, 256 * world), (1, 2048 * world)]:
    g = torch.Generator().manual_seed(S)
    ids = torch.randint(5, 30000, (B, S), generator=g, dtype=torch.int32).to(dev)
    lens = torch.tensor([S, max(200, S - 300)][:B], dtype=torch.int32, device=dev)
    want = model.encode(ids, lens).float()                     # single GPU, whole sequence
    enc = CP.CPT5Encoder(model, B, S)
    s_loc = S // world
    mine = enc.encode_local(ids[:, rank * s_loc:(rank + 1) * s_loc].contiguous(), lens).float()
    ref = want[:, rank * s_loc:(rank + 1) * s_loc]
    pos = torch.arange(rank * s_loc, (rank + 1) * s_loc, device=dev)[None, :, None]
    mask = pos < lens[:, None, None]
    err = ((mine - ref) * mask).abs().max().item()
    good = err < 0.15
    flag = torch.tensor([1.0 if good else 0.0], device=dev)
    if ctx.is_dist:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    allok &= bool(flag.item())

    def t(fn, n=5):
        fn()
        torch.cuda.synchronize()
        if ctx.is_dist:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        x = torch.tensor([a.elapsed_time(b) / n], device=dev)
        if ctx.is_dist:
            dist.all_reduce(x, op=dist.ReduceOp.MAX)
        return float(x.item())

    ms_cp = t(lambda: enc.encode_local(ids[:, rank * s_loc:(rank + 1) * s_loc].contiguous(), lens))
    ms_1 = t(lambda: model.encode(ids, lens))
    if rank == 0:
        print(f"[{'ok' if flag.item() else 'FAIL'}] B={B} S={S} cp={world}: max|err| vs single-GPU encoder {err:.4f};  "
              f"context-parallel {ms_cp:.3f} ms  vs  single GPU {ms_1:.3f} ms  ({ms_1 / ms_cp:.2f}x)", flush=True)
    enc.close()
if rank == 0:
    print("ALL OK" if allok else "SOME FAILED", flush=True)
D.shutdown()
