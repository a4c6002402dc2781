"""End-to-end model numerics + timing on GPU (PYTHONPATH=.)."""
import sys, torch, time
from infomesh_b200.models.bert import BertModel, BGE_SMALL, BGE_RERANKER_BASE, BertConfig
torch.manual_seed(0)
dev = "cuda"; ok_all = True
def check(cfg, B, S, score=False):
    global ok_all
    m = BertModel(cfg, device=dev, seed=1)
    ids = torch.randint(5, cfg.vocab_size, (B, S), device=dev, dtype=torch.int32)
    lens = torch.randint(S // 2, S + 1, (B,), device=dev, dtype=torch.int32)
    if score:
        out = m.score(ids, lens); ref = m.score_ref(ids, lens); tol = 0.02
    else:
        out = m.embed(ids, lens).float(); ref = m.embed_ref(ids, lens); tol = 0.02
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0).item()
    ok = err < tol or cos > 0.999
    ok_all &= ok
    print(dict(model=cfg.name, B=B, S=S, score=score, err=err, cos=cos, ok=ok), flush=True)
    return m
small_cfg = BertConfig(name="tiny", vocab_size=1000, hidden=384, layers=2, heads=12, ffn=1536)
check(small_cfg, 4, 128)
m = check(BGE_SMALL, 8, 128)
check(BGE_SMALL, 3, 40)
r = check(BGE_RERANKER_BASE, 8, 128, score=True)

def timeit(fn, iters=10):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
for (model, B, S, sc) in [(m, 64, 32, False), (m, 1024, 128, False), (r, 1280, 128, True), (r, 256, 512, True)]:
    cfg = model.cfg
    ids = torch.randint(5, cfg.vocab_size, (B, S), device=dev, dtype=torch.int32)
    lens = torch.full((B,), S, device=dev, dtype=torch.int32)
    f = (lambda: model.score(ids, lens)) if sc else (lambda: model.embed(ids, lens))
    g = (lambda: model.score_torch(ids, lens)) if sc else (lambda: model.embed_torch(ids, lens))
    with torch.no_grad():
        t = timeit(f); tt = timeit(g)
    fl = model.flops_per_token(S) * B * S
    print(dict(model=cfg.name, B=B, S=S, ms=t, tflops=fl / t / 1e9, torch_ms=tt, torch_tflops=fl / tt / 1e9), flush=True)
print("MODEL_ALL_OK" if ok_all else "MODEL_FAIL")
