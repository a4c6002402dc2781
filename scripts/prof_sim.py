import torch, sys
from infomesh_b200.ops.search import sim_topk
nq, n = int(sys.argv[1]), int(sys.argv[2])
q = torch.nn.functional.normalize(torch.randn(nq, 384, device="cuda"), dim=1).bfloat16()
d = torch.nn.functional.normalize(torch.randn(n, 384, device="cuda"), dim=1).bfloat16()
for _ in range(2):
    sim_topk(q, d, 10)
torch.cuda.synchronize()
