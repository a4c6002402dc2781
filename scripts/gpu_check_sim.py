"""GPU sanity + timing for sim_topk (run under gpurun with PYTHONPATH=.)."""
import sys, torch
from infomesh_b200.ops.search import sim_topk, sim_topk_ref, sim_topk_partials, topk_merge
torch.manual_seed(0)
dev = "cuda"
ok_all = True
def check(nq, n, dim, k, alive_frac=None):
    global ok_all
    q = torch.nn.functional.normalize(torch.randn(nq, dim, device=dev), dim=1).bfloat16()
    d = torch.nn.functional.normalize(torch.randn(n, dim, device=dev), dim=1).bfloat16()
    alive = None
    if alive_frac is not None:
        alive = (torch.rand(n, device=dev) < alive_frac).to(torch.uint8)
    s, i = sim_topk(q, d, k, alive=alive)
    torch.cuda.synchronize()
    rs, ri = sim_topk_ref(q, d, k, alive=alive)
    kk = rs.shape[1]
    # compare scores (ids may differ on exact ties)
    err = (s[:, :kk] - rs).abs().max().item()
    same = (i[:, :kk] == ri).float().mean().item()
    # score at returned ids must match reference scores
    full = q.float() @ d.float().t()
    gathered = torch.gather(full, 1, i[:, :kk].clamp(min=0))
    err2 = (gathered - s[:, :kk]).abs().max().item()
    ok = err < 2e-3 and err2 < 2e-3 and same > 0.98
    ok_all &= ok
    print(dict(nq=nq, n=n, dim=dim, k=k, alive=alive_frac, err=err, err2=err2, id_match=same, ok=ok), flush=True)
for a in [(1, 1000, 384, 10), (7, 5000, 384, 10), (64, 100000, 384, 10), (128, 300001, 384, 16), (33, 70000, 384, 32),
          (64, 50000, 384, 10, 0.5), (5, 130, 128, 10), (16, 20000, 512, 10)]:
    check(*a)
def bench(nq, n, dim=384, k=10, iters=10):
    q = torch.nn.functional.normalize(torch.randn(nq, dim, device=dev), dim=1).bfloat16()
    d = torch.nn.functional.normalize(torch.randn(n, dim, device=dev), dim=1).bfloat16()
    for _ in range(3): sim_topk(q, d, k)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): sim_topk(q, d, k)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    gbs = n * dim * 2 / ms / 1e6
    # torch baseline: matmul + topk
    for _ in range(2): torch.topk(q @ d.t(), k, dim=1)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): torch.topk(q @ d.t(), k, dim=1)
    e.record(); torch.cuda.synchronize()
    ms_ref = s.elapsed_time(e) / iters
    print(dict(bench_nq=nq, n=n, ms=ms, GBs=gbs, torch_ms=ms_ref), flush=True)
if ok_all:
    for nq in (1, 64, 128):
        bench(nq, 10_000_000)
print("SIM_ALL_OK" if ok_all else "SIM_FAIL")
