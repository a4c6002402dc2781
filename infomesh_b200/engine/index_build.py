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

from infomesh_b200.models.bert import BGE_SMALL, BertModel
from infomesh_b200.ops import dedup as DD
from infomesh_b200.parallel import dist as D


class IndexBuilder:
    """No host synchronisation per batch: the shard's fill level, the seen / duplicate counters and the keep decisions all
    live on the device (``dedup_resolve`` + ``dedup_append`` kernels); the only collective traffic is two push all-gathers
    through the symmetric heap (fingerprints out, per-rank scan results back)."""

    def __init__(self, capacity_per_rank: int, batch_per_rank: int, *, encoder: BertModel | None = None, device=None,
                 threshold: int = DD.HAMMING_THRESHOLD, use_p2p: bool = True):
        self.ctx = D.ctx()
        c = self.ctx
        self.device = torch.device(device) if device is not None else c.device
        self.encoder = encoder or BertModel(BGE_SMALL, device=self.device, seed=1)
        self.threshold, self.bpr = threshold, batch_per_rank
        assert batch_per_rank * c.world <= 8192 and batch_per_rank <= 8192, "use batches of at most 8192 passages across all ranks"
        H = self.encoder.cfg.hidden
        self.vectors = torch.empty((capacity_per_rank, H), device=self.device, dtype=torch.bfloat16)
        self.fingerprints = torch.zeros((capacity_per_rank + 1,), device=self.device, dtype=torch.int64)[:capacity_per_rank]
        self.doc_ids = torch.empty((capacity_per_rank,), device=self.device, dtype=torch.int64)
        self.n_dev = torch.zeros((1,), device=self.device, dtype=torch.int64)        # rows used in this rank's shard
        self.counters = torch.zeros((3,), device=self.device, dtype=torch.int64)     # seen, duplicates, overflow
        self._ag = self._ag_best = None
        if c.is_dist and use_p2p and self.device.type == "cuda":
            from infomesh_b200.parallel import symm

            n_all = c.world * batch_per_rank
            self._heap = symm.SymmetricHeap((4 << 20) + 2 * c.world * (batch_per_rank + n_all) * 8 * 2, c)
            self._ag = symm.AllGatherChannel(self._heap, ((batch_per_rank + 1) // 2 * 2,), torch.int64)
            self._ag_best = symm.AllGatherChannel(self._heap, ((n_all + 1) // 2 * 2,), torch.int64)
            self._stage = torch.zeros(self._ag.shape, device=self.device, dtype=torch.int64)
            self._stage_best = torch.full(self._ag_best.shape, -1, device=self.device, dtype=torch.int64)
            self._heap.barrier()

    @property
    def n(self) -> int:
        return int(self.n_dev.item())

    # ------------------------------------------------------------------ collectives
    def _gather_fp(self, fp: torch.Tensor) -> torch.Tensor:
        """[bpr] local fingerprints -> [world * bpr] in rank order."""
        c = self.ctx
        if not c.is_dist:
            return fp
        if self._ag is not None:
            self._stage[:fp.numel()].copy_(fp)
            return self._ag(self._stage)[:, :fp.numel()].reshape(-1).contiguous()
        return D.all_gather_cat(fp.contiguous()).reshape(-1)

    def _gather_best(self, best: torch.Tensor) -> torch.Tensor:
        """[n] this rank's scan results -> [world, n] (every rank's view of every passage)."""
        c = self.ctx
        if not c.is_dist:
            return best[None]
        if self._ag_best is not None:
            self._stage_best[:best.numel()].copy_(best)
            return self._ag_best(self._stage_best)[:, :best.numel()].contiguous()
        return D.all_gather_cat(best.contiguous())

    # ------------------------------------------------------------------ one batch
    def add_batch(self, token_ids: torch.Tensor, lengths: torch.Tensor, text, word_start, word_end, doc_word_off,
                  first_doc_id: int) -> torch.Tensor:
        """``token_ids`` int32 [bpr, S] / ``lengths`` [bpr]: this rank's passages for the encoder; ``text`` ...
        ``doc_word_off``: the same passages in the normalised-array form of ``ops.dedup.normalize_batch`` (device
        tensors).  Global ids are ``first_doc_id + rank * bpr + i``.  Returns the uint8 keep-mask of this rank's passages
        (0 = near-duplicate of something already indexed or earlier in this batch) -- a device tensor, nothing is synced."""
        c = self.ctx
        bpr = token_ids.shape[0]
        emb = self.encoder.embed(token_ids, lengths)                                       # [bpr, H] bf16, L2-normalised
        fp = DD.simhash_from_arrays(text, word_start, word_end, doc_word_off)              # [bpr] int64 bit patterns
        all_fp = self._gather_fp(fp)                                                       # [world * bpr]
        # (a) each rank scans ITS shard (up to the device-side fill level) for every new passage of every rank
        best = DD.hamming_scan(self.fingerprints, all_fp, self.threshold, n_table_dev=self.n_dev)
        # (b) one kernel: min over the ranks' scan results + the all-pairs "earlier passage of this batch" test
        keep_all = DD.dedup_resolve(all_fp, self._gather_best(best), self.threshold)
        mine = keep_all[c.rank * bpr:(c.rank + 1) * bpr]
        DD.dedup_append(mine, emb, fp, first_doc_id + c.rank * bpr, self.vectors, self.fingerprints, self.doc_ids, self.n_dev,
                        self.counters)
        return mine

    def stats(self) -> dict:
        seen, dups, overflow = (int(x) for x in self.counters.tolist())
        return {"indexed": self.n, "seen": seen, "duplicates": dups, "overflow": overflow, "world": self.ctx.world}
