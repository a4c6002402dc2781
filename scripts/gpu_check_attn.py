"""GPU sanity for attention + nn kernels (PYTHONPATH=.)."""
import sys, torch, math
from infomesh_b200.ops.attention import attention, attention_ref
from infomesh_b200.ops import nn as N
torch.manual_seed(0)
dev = "cuda"; ok_all = True
def check_attn(B, nH, hd, Sq, Sk, lens=False, causal=False, bias=False, scale=None, packed=True):
    global ok_all
    HH = nH * hd
    if packed and Sq == Sk:
        qkv = (torch.randn(B, Sq, 3 * HH, device=dev) * 0.7).bfloat16()
        q, k, v = qkv[..., :HH], qkv[..., HH:2 * HH], qkv[..., 2 * HH:]
    else:
        q = (torch.randn(B, Sq, HH, device=dev) * 0.7).bfloat16()
        kv = (torch.randn(B, Sk, 2 * HH, device=dev) * 0.7).bfloat16()
        k, v = kv[..., :HH], kv[..., HH:]
    kv_lens = torch.randint(1, Sk + 1, (B,), device=dev, dtype=torch.int32) if lens else None
    rb = (torch.randn(nH, Sq + Sk - 1, device=dev) * 0.5) if bias else None
    try:
        o = attention(q, k, v, nH, kv_lens, causal, Sk - Sq if causal else 0, scale, rb)
        torch.cuda.synchronize()
    except Exception as e:
        print("EXC", dict(B=B, nH=nH, hd=hd, Sq=Sq, Sk=Sk), repr(e)[:300], flush=True); ok_all = False; return
    ref = attention_ref(q, k, v, nH, kv_lens, causal, Sk - Sq if causal else 0, scale, rb)
    err = (o.float() - ref).abs().max().item()
    ok = err < 0.03
    ok_all &= ok
    print(dict(B=B, nH=nH, hd=hd, Sq=Sq, Sk=Sk, lens=lens, causal=causal, bias=bias, err=err, ok=ok), flush=True)
for a in [(1, 1, 64, 128, 128), (2, 12, 64, 128, 128), (2, 12, 64, 256, 256, True), (3, 12, 32, 128, 128), (2, 12, 32, 512, 512, True),
          (2, 8, 64, 256, 256, False, True), (2, 8, 64, 256, 256, True, False, True, 1.0), (2, 8, 64, 128, 384, True, False, False, None, False),
          (2, 12, 64, 100, 100, True), (1, 12, 32, 40, 40)]:
    check_attn(*a)

def close(a, b, tol, name):
    global ok_all
    err = (a.float() - b.float()).abs().max().item(); ok = err < tol; ok_all &= ok
    print(dict(test=name, err=err, ok=ok), flush=True)
for H in (384, 768, 512):
    V, S, B = 1000, 64, 4
    ids = torch.randint(0, V, (B * S,), device=dev, dtype=torch.int32)
    word = torch.randn(V, H, device=dev).bfloat16(); pos = torch.randn(512, H, device=dev).bfloat16(); typ = torch.randn(2, H, device=dev).bfloat16()
    g = torch.rand(H, device=dev) + 0.5; b = torch.randn(H, device=dev)
    close(N.embed_ln(ids, word, pos, typ, g, b, 1e-12, S), N.embed_ln_ref(ids, word, pos, typ, g, b, 1e-12, S), 0.06, f"embed_ln{H}")
    x = torch.randn(B * S, H, device=dev).bfloat16(); r = torch.randn(B * S, H, device=dev).bfloat16()
    close(N.layernorm(x, g, b, 1e-12, residual=r), N.layernorm_ref(x, g, b, 1e-12, r), 0.06, f"ln{H}")
    close(N.layernorm(x, g, None, 1e-6, rms_only=True), N.layernorm_ref(x, g, None, 1e-6, None, True), 0.06, f"rms{H}")
    h = torch.randn(B, S, H, device=dev).bfloat16(); lens = torch.tensor([64, 10, 33, 1], device=dev, dtype=torch.int32)
    close(N.pool_norm(h, lens, "cls"), N.pool_norm_ref(h, lens, "cls"), 0.01, f"pool_cls{H}")
    close(N.pool_norm(h, lens, "mean"), N.pool_norm_ref(h, lens, "mean"), 0.01, f"pool_mean{H}")
    w1 = (torch.randn(H, H, device=dev) / math.sqrt(H)).bfloat16(); b1 = torch.randn(H, device=dev) * 0.1
    w2 = (torch.randn(1, H, device=dev) / math.sqrt(H)).bfloat16(); b2 = torch.randn(1, device=dev)
    ref = (torch.tanh(h[:, 0].float() @ w1.float().t() + b1) @ w2.float().t()).squeeze(1) + b2
    close(N.cls_head(h, w1, b1, w2, b2), ref, 0.02, f"cls_head{H}")
lg = torch.randn(7, 32128, device=dev)
v, i = N.row_argmax(lg); torch.cuda.synchronize()
ok = bool((i.long() == lg.argmax(1)).all()); ok_all &= ok; print(dict(test="argmax", ok=ok))
print("ATTN_NN_ALL_OK" if ok_all else "ATTN_NN_FAIL")
