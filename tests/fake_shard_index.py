"""CPU stand-in for GpuSearchIndex inside MultiGpuSearchIndex workers (tests/test_multigpu_front.py): same collective
contract -- every rank ranks the WHOLE corpus identically, and fills doc ids / passage spans only for rows it owns."""
import os

import numpy as np

K_OUT = 5


class FakeShardIndex:
    def __init__(self, store_path, rank, world, kwargs):
        from infomesh_b200.index.local_store import LocalStore

        self.store, self.rank, self.world = LocalStore(store_path), rank, world
        self.nq = int(kwargs.get("query_batch", 8))
        self.fail_on = kwargs.get("fail_on")

    def rebuild(self):
        docs = list(self.store.iter_documents())
        self.all_text = [f"{d.title}\n{d.text}".lower().split() for d in docs]
        per = (len(docs) + self.world - 1) // self.world
        self.lo, self.hi = self.rank * per, min(len(docs), (self.rank + 1) * per)
        self.ids = np.array([int(d.doc_id) for d in docs[self.lo:self.hi]], dtype=np.int64)
        self.text = [d.text for d in docs[self.lo:self.hi]]
        return int(self.ids.size)

    def search_arrays(self, chunk):
        if self.fail_on is not None and self.fail_on == self.rank and any("crash" in q for q in chunk):
            os._exit(3)
        nq = len(chunk)
        scores = np.zeros((nq, K_OUT), np.float32)
        rows = np.full((nq, K_OUT), -1, np.int64)
        ids = np.full((nq, K_OUT), -1, np.int64)
        best = np.full((nq, K_OUT), -1, np.int64)
        span = np.full((nq, K_OUT, 2), -1, np.int64)
        for i, q in enumerate(chunk):
            terms = q.lower().split()
            sc = np.array([sum(t.count(w) for w in terms) for t in self.all_text], dtype=np.float32)
            order = np.argsort(-sc, kind="stable")[:K_OUT]
            for j, r in enumerate(order):
                if sc[r] <= 0:
                    continue
                scores[i, j], rows[i, j] = sc[r], r
                if self.lo <= r < self.hi:
                    ids[i, j] = self.ids[r - self.lo]
                    at = self.text[r - self.lo].lower().find(terms[0])
                    if at >= 0:
                        best[i, j], span[i, j] = 0, (at, at + len(terms[0]))
        return {"scores": scores, "rows": rows, "doc_ids": ids, "pass": best, "span": span}

    def mark_deleted(self, doc_id):
        hit = np.nonzero(self.ids == int(doc_id))[0]
        if hit.size == 0:
            return False
        self.all_text[self.lo + int(hit[0])] = []          # no term matches any more
        return True

    def save(self, directory):
        import json
        from pathlib import Path

        d = Path(directory)
        d.mkdir(parents=True, exist_ok=True)
        blob = json.dumps({"all_text": self.all_text, "lo": self.lo, "hi": self.hi, "ids": self.ids.tolist(), "text": self.text})
        (d / "fake.json").write_text(blob)
        return {"n_docs": int(self.ids.size), "model": "fake", "files": {"fake.json": {"bytes": len(blob)}}}

    def load(self, directory):
        import json
        from pathlib import Path

        st = json.loads((Path(directory) / "fake.json").read_text())
        self.all_text, self.lo, self.hi, self.text = st["all_text"], st["lo"], st["hi"], st["text"]
        self.ids = np.array(st["ids"], dtype=np.int64)
        return int(self.ids.size)

    def stats(self):
        return {"documents": int(self.ids.size), "hbm_bytes": 1000 + self.rank, "rank": self.rank}

    def close(self):
        self.store.close()


def make(store_path, rank, world, kwargs):
    return FakeShardIndex(store_path, rank, world, kwargs)
