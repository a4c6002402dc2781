"""2+ GPU check of the fused tensor-parallel encoder: numerics vs the single-GPU model, CUDA-graph replay, and
timing fused (peer-memory kernels) vs NCCL collectives.  Run under torchrun."""
import sys
from dataclasses import replace

import torch
import torch.distributed as dist

from infomesh_b200.models.bert import BGE_RERANKER_BASE, BertModel
from infomesh_b200.parallel import dist as D
from infomesh_b200.parallel.tp import TPBertModel


def main():
    c = D.init()
    dev = c.device
    layers = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    B, S = (int(sys.argv[2]) if len(sys.argv) > 2 else 64 * c.world), 128
    cfg = replace(BGE_RERANKER_BASE, layers=layers)
    g = torch.Generator(device="cpu").manual_seed(7)
    ids = torch.randint(5, 5000, (B, S), generator=g, dtype=torch.int32).to(dev)
    lens = torch.randint(40, S + 1, (B,), generator=g, dtype=torch.int32).to(dev)
    ref_model = BertModel(cfg, device=dev, seed=11)
    ref_h = ref_model.hidden_states(ids, lens).float()
    ref_logit = ref_model.score(ids, lens)
    ok = True
    results = {}
    for comm in ("fused", "nccl"):
        tpm = TPBertModel(cfg, B, S, seed=11, comm=comm)
        h = tpm.hidden_states(ids, lens).float()
        valid = (torch.arange(S, device=dev)[None] < lens[:, None])[..., None]
        err = ((h - ref_h) * valid).abs().max().item()
        logit = tpm.score(ids, lens)
        lerr = (logit - ref_logit).abs().max().item()
        good = err < 0.15 and lerr < 0.05
        ok &= good
        # second pass must give the same answer (self-advancing channel state)
        h2 = tpm.hidden_states(ids, lens).float()
        rep = ((h2 - h) * valid).abs().max().item()
        ok &= rep == 0.0
        # timing
        for _ in range(3):
            tpm.score(ids, lens)
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n = 10
        for _ in range(n):
            tpm.score(ids, lens)
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / n], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        graph_ms = None
        if comm == "fused":
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                tpm.score(ids, lens)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                out = tpm.score(ids, lens)
            for _ in range(3):
                gr.replay()
            torch.cuda.synchronize()
            gerr = (out - ref_logit).abs().max().item()
            ok &= gerr < 0.05
            dist.barrier()
            a.record()
            for _ in range(n):
                gr.replay()
            b.record()
            torch.cuda.synchronize()
            gm = torch.tensor([a.elapsed_time(b) / n], device=dev)
            dist.all_reduce(gm, op=dist.ReduceOp.MAX)
            graph_ms = gm.item()
        results[comm] = (err, lerr, rep, ms.item(), graph_ms)
        if c.rank == 0:
            print(f"{comm:5s}: hidden err {err:.4f}  logit err {lerr:.4f}  repeat diff {rep:.1e}  {ms.item():.3f} ms/forward"
                  + (f"  graph {graph_ms:.3f} ms" if graph_ms else "") + ("  ok" if good else "  MISMATCH"))
    # single-GPU time of the same model for reference
    for _ in range(3):
        ref_model.score(ids, lens)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ref_model.score(ids, lens)
    b.record()
    torch.cuda.synchronize()
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if c.rank == 0:
        print(f"single GPU: {a.elapsed_time(b) / 10:.3f} ms/forward   (B={B}, S={S}, layers={layers}, tp={c.world})")
        print("ALL OK" if flag.item() == 1.0 else "FAILED")
    D.shutdown()


if __name__ == "__main__":
    main()
