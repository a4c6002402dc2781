import torch
from infomesh_b200 import _native
from infomesh_b200.ops import mx as MX
_native.require()
DEV="cuda"; m,n,k=90112,768,768
a = MX.quantize_act_ref(torch.randn(m, k, device=DEV)); w = MX.quantize_weight(torch.randn(n, k, device=DEV) * 0.05)
bias = torch.randn(n, device=DEV); res = torch.randn(m, n, device=DEV).bfloat16(); g = torch.rand(n, device=DEV)+0.5; b = torch.randn(n, device=DEV)
mxo = MX.alloc_act(m, n, DEV); out = torch.empty((m, n), device=DEV, dtype=torch.bfloat16)
for _ in range(3): MX.linear_mx_ln(a, w, bias, res, g, b, 1e-5, mxo, out=out)
torch.cuda.synchronize()
