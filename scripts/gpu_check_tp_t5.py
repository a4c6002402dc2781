#!/usr/bin/env python
"""Tensor-parallel T5 (parallel/tp_t5.py) against the single-GPU model, and a latency table fused vs NCCL vs 1 GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 scripts/gpu_check_tp_t5.py [t5-small|t5-base]
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from infomesh_b200.models.t5 import CONFIGS, T5Model  # noqa: E402
from infomesh_b200.parallel import dist as D  # noqa: E402
from infomesh_b200.parallel.tp_t5 import TPT5Model  # noqa: E402

ctx = D.init()
rank, world, dev = ctx.rank, ctx.world, ctx.device
cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "t5-small"]
S, T = 128, 24
allok = True


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    if ctx.is_dist:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    if ctx.is_dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


rows = []
single = T5Model(cfg, device=dev, seed=7)
NVLS = os.environ.get("TP_T5_NVLS", "1") == "1"
GRAPH = os.environ.get("TP_T5_GRAPH", "1") == "1"
for B in [int(x) for x in os.environ.get("TP_T5_BATCHES", "1,8,64").split(",")]:
    g = torch.Generator().manual_seed(B)
    ids = torch.randint(5, cfg.vocab_size, (B, S), generator=g, dtype=torch.int32).to(dev)
    lens = torch.randint(S // 2, S + 1, (B,), generator=g, dtype=torch.int32).to(dev)
    want_enc = single.encode(ids, lens).float()
    want_tok = single.generate(ids, lens, max_new_tokens=T, check_every=10 ** 6)
    res = {}
    for comm in ("fused", "nccl"):
        m = TPT5Model(cfg, B, S, seed=7, comm=comm, prefer_nvls=NVLS)
        enc = m.encode(ids, lens).float()
        mask = (torch.arange(S, device=dev)[None] < lens[:, None])[..., None]
        err = ((enc - want_enc) * mask).abs().max().item()
        tok = m.generate(ids, lens, max_new_tokens=T, use_graph=GRAPH)
        agree = (tok == want_tok[:, :T]).float().mean().item()
        same = torch.tensor([int(tok.sum().item())], device=dev)
        lo, hi = same.clone(), same.clone()
        if ctx.is_dist:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        good = err < 0.25 and agree > 0.85 and int(lo.item()) == int(hi.item())
        allok &= good
        t_enc = timeit(lambda: m.encode(ids, lens))
        t_gen = timeit(lambda: m.generate(ids, lens, max_new_tokens=T, use_graph=GRAPH), iters=3)
        res[comm] = (t_enc, t_gen)
        if rank == 0:
            print(f"[{'ok' if good else 'FAIL'}] tp={world} B={B} {comm:5s} nvls={getattr(m, 'nvls', False)}: max|enc err| {err:.3f}, token agreement with 1 GPU "
                  f"{agree:.3f}, identical on all ranks {int(lo.item()) == int(hi.item())}", flush=True)
        m.close()
    t1_enc = timeit(lambda: single.encode(ids, lens))
    t1_gen = timeit(lambda: single.generate(ids, lens, max_new_tokens=T, check_every=10 ** 6), iters=3)
    rows.append((B, t1_enc, res["fused"][0], res["nccl"][0], t1_gen, res["fused"][1], res["nccl"][1]))
if rank == 0:
    print(f"\n{cfg.name}, S={S}, {T} new tokens, tp={world} (ms, device time, max over ranks)")
    print("| B | encode 1 GPU | encode TP fused | encode TP NCCL | generate 1 GPU | generate TP fused | generate TP NCCL | fused vs NCCL (generate) |")
    print("|---|---|---|---|---|---|---|---|")
    for B, a, b, c, d, e, f in rows:
        print(f"| {B} | {a:.3f} | {b:.3f} | {c:.3f} | {d:.2f} | {e:.2f} | {f:.2f} | {f / e:.2f}x |")
    print("ALL OK" if allok else "SOME FAILED", flush=True)
D.shutdown()
