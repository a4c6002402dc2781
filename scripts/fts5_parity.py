"""BM25 ordering parity against SQLite FTS5 at scale (reference infomesh/index/local_store.py:316-332).

    python scripts/fts5_parity.py [--docs 1000000] [--queries 300] [--out profiles/fts5_parity.json]

Builds ONE synthetic corpus (Zipf vocabulary, title + body) into the product's ``LocalStore`` (FTS5, ``bm25()`` order) and
into the posting builder that feeds the device index, then compares the top-10 of every query: the CUDA BM25 kernel when a
GPU is visible, the NumPy oracle of the same formula otherwise.  A rank agrees when the document is the same, or when the
two scores are equal to 1e-5 relative (FTS5 leaves the order of exact ties unspecified).  Exit code 1 below 99.5 %."""
import argparse
import json
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=300)
    ap.add_argument("--vocab", type=int, default=60_000)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch

    from infomesh_b200.index.local_store import LocalStore
    from infomesh_b200.ops import bm25 as BM

    rng = np.random.default_rng(11)
    p = 1.0 / np.arange(1, a.vocab + 1) ** 1.07
    p /= p.sum()
    cdf = np.cumsum(p)
    words = np.array([f"t{i}x" for i in range(a.vocab)])
    tmp = tempfile.mkdtemp(prefix="im_fts_")
    st = LocalStore(Path(tmp) / "index.db")
    b = BM.HostIndexBuilder()
    t0 = time.time()
    CH = 20000
    for base in range(0, a.docs, CH):
        n = min(CH, a.docs - base)
        lens = rng.integers(12, 60, size=n)
        ids = np.searchsorted(cdf, rng.random(int(lens.sum())))
        pos = 0
        st._conn.execute("BEGIN")
        for i in range(n):
            w = words[ids[pos:pos + lens[i]]]
            pos += lens[i]
            title, text = " ".join(w[:3]), " ".join(w[3:])
            k = base + i
            st._conn.execute("INSERT INTO documents (url, title, text, raw_html_hash, text_hash, crawled_at) VALUES (?,?,?,?,?,?)",
                             (f"https://h{k % 997}.example/{k}", title, text, f"r{k}", f"x{k}", 1.7e9))
            b.add_text(f"{title}\n{text}")
        st._conn.execute("COMMIT")
        if base % 200000 == 0:
            print(f"  {base + n} docs, {time.time() - t0:.0f}s", flush=True)
    t_build = time.time() - t0
    csr = b.export()
    cuda = torch.cuda.is_available()
    index = BM.Bm25Index(csr, device="cuda") if cuda else None
    # query mix: 1-3 terms; a third drawn from the head of the distribution (df up to ~10 % of N)
    queries = []
    for qi in range(a.queries):
        nt = 1 + qi % 3
        if qi % 3 == 0:
            t = rng.integers(0, 50, size=nt)
        else:
            t = np.searchsorted(cdf, rng.random(nt))
        queries.append(list(dict.fromkeys(words[t].tolist())))
    agree = ranks = exact_lists = nonempty = 0
    worst = 0.0
    t_fts = t_ours = 0.0
    for q in queries:
        t0 = time.time()
        fts = st.search(" ".join(q), limit=10)
        t_fts += time.time() - t0
        terms = [b.lookup(w) for w in q]
        t0 = time.time()
        if cuda:
            tt = torch.full((1, 8), -1, dtype=torch.int32, device="cuda")
            tt[0, :len(terms)] = torch.tensor(terms, dtype=torch.int32)
            s, d = index.search(tt, k=10)
            ours = [(float(x), int(y)) for x, y in zip(s[0].cpu(), d[0].cpu()) if y >= 0]
        else:
            ours = BM.bm25_ref(csr, terms, k=10)
        t_ours += time.time() - t0
        f = [(float(r.score), int(r.doc_id) - 1) for r in fts]
        nonempty += bool(f)
        if len(f) != len(ours):
            ranks += max(len(f), len(ours))
            continue
        same = True
        for (sf, df_), (so, do) in zip(f, ours):
            ranks += 1
            rel = abs(sf - so) / max(abs(sf), 1e-9)
            worst = max(worst, rel)
            if df_ == do or rel < 1e-5:
                agree += 1
            same &= df_ == do
        exact_lists += same
    res = {"docs": a.docs, "queries": len(queries), "non_empty": nonempty, "backend": "cuda" if cuda else "numpy-oracle",
           "rank_agreement": round(agree / max(ranks, 1), 5), "identical_lists": exact_lists, "worst_rel_score_err": worst,
           "build_s": round(t_build, 1), "fts5_ms_per_query": round(1e3 * t_fts / len(queries), 3),
           "ours_ms_per_query_b1": round(1e3 * t_ours / len(queries), 3), "vocab": a.vocab, "nnz": int(csr["off"][-1])}
    print(json.dumps(res))
    if a.out:
        Path(a.out).write_text(json.dumps(res, indent=1) + "\n")
    sys.exit(0 if res["rank_agreement"] >= 0.995 and nonempty > len(queries) // 2 else 1)


if __name__ == "__main__":
    main()
