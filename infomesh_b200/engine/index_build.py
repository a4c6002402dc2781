"""Index-build pipeline (BASELINE config 5: "encode + SimHash-dedup N passages"), data-parallel over the ranks.

Per batch and rank: the encoder embeds the rank's slice of the passages, the SimHash kernel fingerprints them, the new
fingerprints are all-gathered (pushed straight into every peer's buffer through the symmetric heap when one is
available, NCCL / gloo otherwise) so that every rank can scan ITS shard of the global fingerprint table for every new
passage, the per-rank nearest matches are min-reduced, and passages without a near-duplicate (Hamming distance <= 3)
are appended to the owning rank's vector / fingerprint shard.

Replaces the reference's sequential ``add_document`` loop with a per-document Python SimHash and dict scan
(infomesh/crawler/dedup.py:167-181, infomesh/crawler/simhash.py:43-109, infomesh/index/vector_store.py:124-173)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from infomesh_b200.models.bert import BGE_SMALL, BertModel
from infomesh_b200.ops import dedup as DD
from infomesh_b200.parallel import dist as D


class IndexBuilder:
    def __init__(self, capacity_per_rank: int, batch_per_rank: int, *, encoder: BertModel | None = None, device=None,
                 threshold: int = DD.HAMMING_THRESHOLD, use_p2p: bool = True):
        self.ctx = D.ctx()
        c = self.ctx
        self.device = torch.device(device) if device is not None else c.device
        self.encoder = encoder or BertModel(BGE_SMALL, device=self.device, seed=1)
        self.threshold, self.bpr = threshold, batch_per_rank
        H = self.encoder.cfg.hidden
        self.vectors = torch.empty((capacity_per_rank, H), device=self.device, dtype=torch.bfloat16)
        self.fingerprints = torch.empty((capacity_per_rank,), device=self.device, dtype=torch.int64)
        self.doc_ids = torch.empty((capacity_per_rank,), device=self.device, dtype=torch.int64)
        self.n = 0                      # rows used in this rank's shard
        self.seen = self.dups = 0
        self._ag = None
        if c.is_dist and use_p2p and self.device.type == "cuda":
            from infomesh_b200.parallel import symm

            self._heap = symm.SymmetricHeap(4 << 20, c)
            self._ag = symm.AllGatherChannel(self._heap, ((batch_per_rank + 1) // 2 * 2,), torch.int64)
            self._stage = torch.zeros(self._ag.shape, device=self.device, dtype=torch.int64)
            self._heap.barrier()

    # ------------------------------------------------------------------ collectives
    def _gather_fp(self, fp: torch.Tensor) -> torch.Tensor:
        """[bpr] local fingerprints -> [world * bpr] in rank order."""
        c = self.ctx
        if not c.is_dist:
            return fp
        if self._ag is not None:
            self._stage[:fp.numel()].copy_(fp)
            return self._ag(self._stage)[:, :fp.numel()].reshape(-1)
        return D.all_gather_cat(fp.contiguous()).reshape(-1)

    # ------------------------------------------------------------------ one batch
    def add_batch(self, token_ids: torch.Tensor, lengths: torch.Tensor, text, word_start, word_end, doc_word_off,
                  first_doc_id: int) -> torch.Tensor:
        """``token_ids`` int32 [bpr, S] / ``lengths`` [bpr]: this rank's passages for the encoder; ``text`` ...
        ``doc_word_off``: the same passages in the normalised-array form of ``ops.dedup.normalize_batch`` (device
        tensors).  Global ids are ``first_doc_id + rank * bpr + i``.  Returns the boolean keep-mask of this rank's
        passages (False = near-duplicate of something already indexed or earlier in this batch)."""
        c = self.ctx
        bpr = token_ids.shape[0]
        emb = self.encoder.embed(token_ids, lengths)                                       # [bpr, H] bf16, L2-normalised
        fp = DD.simhash_from_arrays(text, word_start, word_end, doc_word_off)              # [bpr] int64 bit patterns
        all_fp = self._gather_fp(fp)                                                       # [world * bpr]
        # (a) against everything already indexed: each rank scans its own shard for every new passage, then min-reduce
        best = DD.hamming_scan(self.fingerprints[:self.n], all_fp, self.threshold)
        if c.is_dist:
            # packed (distance << 32 | index) with all-ones = none: as SIGNED int64 'none' is -1, so reduce the
            # unsigned order by flipping the sign bit around the MIN
            flipped = best ^ torch.iinfo(torch.int64).min
            dist.all_reduce(flipped, op=dist.ReduceOp.MIN)
            best = flipped ^ torch.iinfo(torch.int64).min
        dup_old = best != -1
        # (b) inside the batch: passage j is dropped if an EARLIER passage of the batch is within the threshold
        #     (all-pairs on the gathered fingerprints: world * bpr <= 8192 keeps this a few hundred MB at most)
        assert all_fp.numel() <= 8192, "use batches of at most 8192 passages across all ranks"
        x = all_fp[:, None] ^ all_fp[None, :]
        pop = torch.zeros_like(x, dtype=torch.int32)
        for shift in range(0, 64, 16):                                                     # popcount via 16-bit table-free folding
            pop += _popc16((x >> shift) & 0xFFFF)
        earlier = torch.tril(torch.ones_like(pop, dtype=torch.bool), diagonal=-1)
        dup_new = ((pop <= self.threshold) & earlier).any(dim=1)
        keep_all = ~(dup_old | dup_new)
        mine = keep_all[c.rank * bpr:(c.rank + 1) * bpr]
        k = int(mine.sum().item())
        if self.n + k > self.vectors.shape[0]:
            raise RuntimeError("index shard is full")
        ids = first_doc_id + c.rank * bpr + torch.arange(bpr, device=self.device)
        self.vectors[self.n:self.n + k] = emb[mine]
        self.fingerprints[self.n:self.n + k] = fp[mine]
        self.doc_ids[self.n:self.n + k] = ids[mine]
        self.n += k
        self.seen += bpr
        self.dups += bpr - k
        return mine

    def stats(self) -> dict:
        return {"indexed": self.n, "seen": self.seen, "duplicates": self.dups, "world": self.ctx.world}


def _popc16(v: torch.Tensor) -> torch.Tensor:
    v = v - ((v >> 1) & 0x5555)
    v = (v & 0x3333) + ((v >> 2) & 0x3333)
    v = (v + (v >> 4)) & 0x0F0F
    return ((v + (v >> 8)) & 0x1F).to(torch.int32)
