#!/usr/bin/env python
"""Multi-GPU check of the VMM symmetric heap + NVLS multicast kernels (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/gpu_check_nvls.py

Correctness of mc all-gather / reduce-scatter / all-reduce against NCCL, and a latency table (device time, max over
ranks): NVLS multicast vs the unicast P2P push kernels (parallel/symm.py) vs NCCL."""
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from infomesh_b200.parallel import dist as D  # noqa: E402
from infomesh_b200.parallel import symm, vmm  # noqa: E402

ctx = D.init()
rank, world, dev = ctx.rank, ctx.world, ctx.device
cap = vmm.support(dev.index or 0)
if rank == 0:
    print(f"world={world} capabilities={cap}", flush=True)
heap = vmm.VmmHeap(256 << 20, ctx)
if rank == 0:
    print(f"VMM heap mapped: {heap.nbytes >> 20} MiB per rank, multicast={'yes' if heap.has_multicast else 'NO (P2P only)'}", flush=True)

# peer mapping sanity: write my rank into my slot of every peer through the unicast mappings
probe, off = heap.alloc((world,), torch.int32)
tab = heap.peer_table(off)
torch.cuda.synchronize()
dist.barrier()
for p in range(world):
    t = torch.as_tensor(symm._RawCuda(heap.bases[p] + off, world * 4), device=dev).view(torch.int32)
    t[rank] = rank + 1
torch.cuda.synchronize()
dist.barrier()
ok = bool((probe.cpu() == torch.arange(1, world + 1, dtype=torch.int32)).all())
print(f"[{'ok' if ok else 'FAIL'}] rank {rank}: unicast peer stores through VMM mappings", flush=True)
allok = ok


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


rows = []
if heap.has_multicast:
    pheap = symm.SymmetricHeap(128 << 20, ctx)
    for n_elem in (384 * 8, 64 * 384, 1 << 20):           # 8 / 64 query embeddings (bf16), 2 MiB
        src = (torch.randn(n_elem, device=dev) + rank).bfloat16()
        mc = vmm.McAllGather(heap, (n_elem,), torch.bfloat16)
        pp = symm.AllGatherChannel(pheap, (n_elem,), torch.bfloat16)
        ref = [torch.empty_like(src) for _ in range(world)]
        dist.all_gather(ref, src)
        got = mc(src)
        torch.cuda.synchronize()
        good = all(torch.equal(got[p], ref[p]) for p in range(world))
        allok &= good
        for _ in range(3):                                 # channel reuse (step counter, parity buffers)
            good &= all(torch.equal(mc(src)[p], ref[p]) for p in range(world))
        t_mc, t_p2p = timeit(lambda: mc(src)), timeit(lambda: pp(src))
        t_nccl = timeit(lambda: dist.all_gather(ref, src))
        rows.append(("all-gather", n_elem * 2, t_mc, t_p2p, t_nccl))
        if rank == 0:
            print(f"[{'ok' if good else 'FAIL'}] mc all-gather {n_elem * 2} B/rank", flush=True)
    for shape in ((1024, 768), (8192, 768)):               # GEMM -> reduce-scatter consumer sizes (bf16 partial sums)
        red = vmm.McReduce(heap, shape, torch.bfloat16)
        part = (torch.randn(shape, device=dev) * 0.5).bfloat16()
        full = part.float().clone()
        dist.all_reduce(full)
        red.buf.copy_(part)
        mine = red.reduce_scatter().float()
        torch.cuda.synchronize()
        n = part.numel() // world
        want = full.reshape(-1)[rank * n:(rank + 1) * n]
        err = (mine - want).abs().max().item()
        good = err < 0.02 * want.abs().max().item() + 0.05
        red2 = vmm.McReduce(heap, shape, torch.bfloat16)
        red2.buf.copy_(part)
        ar = red2.all_reduce().float()
        torch.cuda.synchronize()
        err2 = (ar - full).abs().max().item()
        good &= err2 < 0.02 * full.abs().max().item() + 0.05
        allok &= good
        out_rs = torch.empty(n, device=dev, dtype=torch.bfloat16)
        t_mc = timeit(lambda: red.reduce_scatter())
        t_nccl = timeit(lambda: dist.reduce_scatter_tensor(out_rs, part.reshape(-1)))
        t_ar = timeit(lambda: red2.all_reduce())
        t_ar_nccl = timeit(lambda: dist.all_reduce(part))
        rows.append((f"reduce-scatter {shape}", part.numel() * 2, t_mc, float("nan"), t_nccl))
        rows.append((f"all-reduce {shape}", part.numel() * 2, t_ar, float("nan"), t_ar_nccl))
        if rank == 0:
            print(f"[{'ok' if good else 'FAIL'}] mc reduce-scatter / all-reduce {shape}: max err {err:.4g} / {err2:.4g}", flush=True)
    if rank == 0:
        print("\n| op | bytes/rank | NVLS multimem (us) | unicast P2P push (us) | NCCL (us) |\n|---|---|---|---|---|")
        for name, b, a, p2, n_ in rows:
            print(f"| {name} | {b} | {a:.1f} | {p2:.1f} | {n_:.1f} |")
    pheap.close()
if rank == 0:
    print("ALL OK" if allok else "SOME FAILED", flush=True)
heap.close()
D.shutdown()
