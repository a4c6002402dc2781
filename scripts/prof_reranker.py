import torch, sys
from infomesh_b200.models.bert import BertModel, BGE_RERANKER_BASE
B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 128
m = BertModel(BGE_RERANKER_BASE, device="cuda", seed=1)
ids = torch.randint(5, 1000, (B, S), device="cuda", dtype=torch.int32)
lens = torch.full((B,), S, device="cuda", dtype=torch.int32)
for _ in range(2):
    m.score(ids, lens)
torch.cuda.synchronize()
