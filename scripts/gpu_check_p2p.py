"""2+ GPU check of the peer-memory backend: heap mapping, p2p all-gather vs NCCL, fused top-k exchange vs oracle,
and graph-captured multi-step replay.  Run under torchrun."""
import torch
import torch.distributed as dist

from infomesh_b200.ops import search as S
from infomesh_b200.parallel import dist as D
from infomesh_b200.parallel import symm


def main():
    c = D.init()
    dev = c.device
    heap = symm.SymmetricHeap(16 << 20)
    ok = True
    # ---- all-gather
    ag = symm.AllGatherChannel(heap, (4096,), torch.float32)
    for step in range(5):
        src = torch.full((4096,), float(c.rank * 100 + step), device=dev) + torch.arange(4096, device=dev) * 1e-3
        got = ag(src).clone()
        ref = D.all_gather_cat(src)
        good = torch.equal(got, ref)
        ok &= good
        if c.rank == 0:
            print("allgather step", step, "ok" if good else "MISMATCH")
    # ---- fused top-k exchange
    nq, k = 64, 20
    ch = symm.TopkChannel(heap, nq, k)
    g = torch.Generator(device=dev).manual_seed(1234 + c.rank)
    for step in range(4):
        sc = torch.rand((3, nq, k), device=dev, generator=g).sort(dim=2, descending=True).values
        ids = torch.randint(0, 1 << 30, (3, nq, k), device=dev, generator=g) * c.world + c.rank
        ls, li = S.topk_merge(sc, ids, k, push=ch)
        gs, gi = S.topk_merge(ch.cand_scores, ch.cand_ids, k, wait=ch)
        all_s, all_i = D.all_gather_cat(ls), D.all_gather_cat(li)
        rs, ri = S.topk_merge(all_s, all_i, k)
        good = torch.equal(gs, rs) and torch.equal(gi, ri)
        ok &= good
        if c.rank == 0:
            print("topk exchange step", step, "ok" if good else "MISMATCH")
    # ---- degraded mode: rank (world-1) never pushes; every rank drops that shard after the poll limit and says so
    chd = symm.TopkChannel(heap, nq, k, degraded_ok=True, wait_limit=200_000)
    sc = torch.rand((2, nq, k), device=dev, generator=g).sort(dim=2, descending=True).values
    ids = torch.randint(0, 1 << 30, (2, nq, k), device=dev, generator=g) * c.world + c.rank
    silent = c.world - 1
    for step in range(2):
        ls, li = S.topk_merge(sc, ids, k, push=None if c.rank == silent else chd)
        gs, gi = S.topk_merge(chd.cand_scores, chd.cand_ids, k, wait=chd)
        all_s, all_i = D.all_gather_cat(ls), D.all_gather_cat(li)
        rs, ri = S.topk_merge(all_s[:silent].contiguous(), all_i[:silent].contiguous(), k)
        good = torch.equal(gs, rs) and torch.equal(gi, ri) and chd.dead_ranks() == [silent]
        ok &= good
        if c.rank == 0:
            print("degraded exchange step", step, "ok" if good else f"MISMATCH dead={chd.dead_ranks()}")
    # ---- graph replay
    src = torch.zeros((4096,), device=dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ag(src)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        out = ag(src)
    for it in range(6):
        src.fill_(float(10 * it + c.rank))
        gr.replay()
        torch.cuda.synchronize()
        want = torch.tensor([10.0 * it + r for r in range(c.world)], device=dev)
        good = torch.equal(out[:, 0], want) and torch.equal(out[:, -1], want)
        ok &= good
    if c.rank == 0:
        print("graph replay", "ok" if ok else "MISMATCH")
    # ---- timing: p2p all-gather vs NCCL for the bench-sized block
    small = torch.zeros((64 * 20,), device=dev)
    agc = symm.AllGatherChannel(heap, (64 * 20,), torch.float32)

    def t(fn, n=200):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e3

    def p2p():
        agc(small)

    us_p2p = t(p2p)
    us_nccl = t(lambda: D.all_gather_cat(small))
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if c.rank == 0:
        print(f"allgather 5 KB x{c.world}: p2p {us_p2p:.1f} us   nccl {us_nccl:.1f} us")
        print("ALL OK" if flag.item() == 1.0 else "FAILED")
    heap.close()
    D.shutdown()


if __name__ == "__main__":
    main()
