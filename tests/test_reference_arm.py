"""The reference arm of bench.py (baseline/): corpus-generator parity with the engine's generator, the stand-in wheels,
and a tiny end-to-end run through the UNMODIFIED reference's search_hybrid()."""
from __future__ import annotations

import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "baseline"))
import reference_arm as R  # noqa: E402


def test_generators_agree_with_engine_synth():
    from infomesh_b200.engine import synth

    cfg = synth.SynthConfig(n_docs=300, n_docs_global=300)
    assert (cfg.vocab_terms, cfg.doc_len, cfg.dim, cfg.zipf_s, cfg.seed) == (R.VOCAB_TERMS, R.DOC_LEN, R.DIM, R.ZIPF_S, R.SEED)
    cdf = R._zipf_cdf()
    assert torch.equal(cdf, synth.zipf_cdf(cfg.vocab_terms, cfg.zipf_s, "cpu"))
    a = R.doc_terms(37, 200, cdf)
    b = synth.gen_doc_terms(cfg, "cpu", 37, 200, cdf)
    assert torch.equal(a.to(torch.int32), b)
    va = R.doc_vectors(5, 64)
    vb = synth.gen_vectors(cfg, "cpu", 5, 64).float()
    assert torch.allclose(va, vb, atol=4e-3)          # engine stores bf16


def test_query_terms_match_engine_queries():
    from infomesh_b200.engine import synth

    cfg = synth.SynthConfig(n_docs=1000, n_docs_global=1000)
    for mix in ("rare", "common"):
        qt, _, _, docs = synth.make_queries(cfg, 12, mix=mix)
        mine, docs2 = R.query_terms(1000, 12, R._zipf_cdf(), mix=mix)
        assert torch.equal(docs, docs2)
        for i, terms in enumerate(mine):
            assert terms == [int(t) for t in qt[i] if int(t) >= 0]


def test_zstandard_stand_in_round_trip():
    sys.path.append(str(ROOT / "baseline" / "shims"))
    import importlib

    z = importlib.import_module("zstandard")
    if "shims" not in (getattr(z, "__file__", "") or ""):
        pytest.skip("real zstandard wheel present")
    blob = z.ZstdCompressor(level=3).compress(b"abc" * 1000)
    assert len(blob) < 200
    assert z.ZstdDecompressor().decompress(blob, max_output_size=10_000) == b"abc" * 1000


def test_install_manifest_matches_source_when_mounted():
    ref = Path("/root/reference/infomesh")
    if not ref.is_dir():
        pytest.skip("reference source not mounted")
    import install_ref

    install_ref.install()
    assert install_ref.verify() == 0


@pytest.mark.timeout(600)
def test_reference_arm_runs_search_hybrid():
    if not (ROOT / "baseline" / "_ref" / "infomesh").is_dir() and not Path("/root/reference/infomesh").is_dir():
        pytest.skip("reference not installed")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--batch", "4", "--ref-docs", "2000"], capture_output=True, text=True, timeout=580, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and "unavailable" not in line, line
    assert line["value"] > 0 and line["config"]["index_docs"] == 2000
    assert line["same_config"] is False and line["gpu_launches"] == 0
    # the 2-3 AND-ed terms come from a real document, so FTS5 contributed to the fused top-10
    assert line["config"]["last_query_fts_hits_in_top10"] >= 1


def test_merkle_roots_and_proofs_interoperate_with_the_reference():
    """Proofs are wire format: a proof built by either implementation must verify in the other (scripts/diff_interop.py
    covers the codec; this covers trust/merkle.py)."""
    import hashlib
    import sys

    root = Path(__file__).resolve().parent.parent / "baseline"
    if not (root / "_ref" / "infomesh" / "trust" / "merkle.py").exists():
        pytest.skip("reference not installed")
    sys.path[:0] = [str(root / "shims"), str(root / "_ref")]
    try:
        from infomesh.trust import merkle as R
    finally:
        del sys.path[:2]
    from infomesh_b200.trust import merkle as M

    for n in (1, 2, 3, 5, 8, 13, 64):
        hs = [hashlib.sha256(str(i).encode()).hexdigest() for i in range(n)]
        a, b = R.MerkleTree(), M.MerkleTree()
        assert a.build(hs) == b.build(hs)
        for i in range(n):
            wire_theirs, wire_ours = R.serialize_proof(a.get_proof(i)), M.serialize_proof(b.get_proof(i))
            assert [(h, str(s)) for h, s in wire_theirs["proof_path"]] == [(h, str(s)) for h, s in wire_ours["proof_path"]]
            assert M.MerkleTree.verify_document(hs[i], M.deserialize_proof(wire_theirs))
            assert R.MerkleTree.verify_document(hs[i], R.deserialize_proof(wire_ours))
