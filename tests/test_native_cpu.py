"""Host-side native code + CPU oracles (no GPU needed): C++ tokenizer / CSR builder / MD5 / SimHash, BM25 oracle,
synthetic corpus determinism, fusion oracles, build entry point."""
from __future__ import annotations

import hashlib

import numpy as np
import pytest
import torch

from infomesh_b200 import _native
from infomesh_b200.ops import bm25 as BM
from infomesh_b200.ops import dedup as DD
from infomesh_b200.ops import fuse as F

native = pytest.mark.skipif(not _native.available(), reason="native library not built")


@native
def test_native_library_exports():
    L = _native.lib()
    assert L.im_abi_version() == 1
    for sym in ("im_gemm_bf16_tn", "im_sim_topk", "im_topk_merge", "im_attn_fwd", "im_attn_decode", "im_embed_ln",
                "im_sum_ln", "im_pool_norm", "im_cls_head", "im_row_argmax", "im_bm25_topk", "im_passage_score",
                "im_simhash", "im_hamming_scan", "im_rrf_fuse", "im_build_pairs", "im_rerank_select"):
        assert hasattr(L, sym), sym


@native
def test_sass_contains_blackwell_tensor_ops():
    """Compile-only evidence that the hot kernels are tcgen05 / TMA, not legacy mma.sync."""
    import shutil
    import subprocess

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", str(_native.lib_path())], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass and "UTMASTG" in sass
    assert "HMMA.16816" not in sass


@native
def test_md5_and_simhash_match_reference_semantics():
    import ctypes

    L = _native.lib()
    L.im_md5_first8_be.restype = ctypes.c_ulonglong
    for s in (b"", b"abc", b"x" * 55, b"y" * 56, b"z" * 64, b"w" * 119, b"q" * 1000):
        assert L.im_md5_first8_be(s, ctypes.c_longlong(len(s))) == int.from_bytes(hashlib.md5(s).digest()[:8], "big")
    texts = ["The quick brown fox jumps over the lazy dog", "", "one", "one two", "Ünïcödé wörds hérè too",
             "under_score and-dash, punctuation!", "a b c d e f g " * 50]
    got = DD.simhash_cpu(texts)
    want = np.asarray([DD.simhash_py(t) for t in texts], dtype=np.uint64)
    assert (got == want).all()
    assert DD.simhash_py("") == 0 and DD.hamming(0b1011, 0b0001) == 2
    a, b = DD.simhash_py("the cat sat on the mat today " * 20), DD.simhash_py("the cat sat on the mat today " * 20 + "x")
    assert DD.hamming(a, b) <= 3


@native
def test_index_builder_and_bm25_oracle():
    b = BM.HostIndexBuilder()
    docs = ["The quick brown fox jumps over the lazy dog", "Quick brown dogs are lazy", "quick quick quick fox",
            "Python asyncio tutorial", "völlig andere Wörter"]
    ids = [b.add_text(d) for d in docs]
    assert ids == [0, 1, 2, 3, 4] and b.n_docs == 5 and b.lookup("QUICK") >= 0 and b.lookup("missing") == -1
    csr = b.export()
    t = b.lookup("quick")
    assert csr["df"][t] == 3 and list(csr["doc"][csr["off"][t]:csr["off"][t + 1]]) == [0, 1, 2]
    assert list(csr["tf"][csr["off"][t]:csr["off"][t + 1]]) == [1, 1, 3]
    q = b.tokenize("quick fox")
    res = BM.bm25_ref(csr, q, k=5)
    assert [d for _, d in res][0] in (0, 2) and {d for _, d in res} == {0, 2}      # implicit AND
    assert BM.bm25_ref(csr, [t, 9999], k=5) == []                                  # unknown term empties the AND
    assert b.term(t) == "quick"
    # parity with SQLite FTS5 ordering on the same tiny corpus
    from infomesh_b200.index.local_store import LocalStore

    s = LocalStore()
    for i, d in enumerate(docs):
        s.add_document(f"https://d/{i}", "", d, f"r{i}", f"t{i}")
    fts = [r.doc_id - 1 for r in s.search("quick fox")]
    assert fts == [d for _, d in res]


def test_synth_corpus_is_shard_independent():
    from infomesh_b200.engine.synth import SynthConfig, SynthShard, make_queries

    full = SynthShard(SynthConfig(n_docs=3000, n_docs_global=3000, vocab_terms=2000, doc_len=24, passage_len=32),
                      device="cpu", build_chunk=1100)
    half = SynthShard(SynthConfig(n_docs=1500, n_docs_global=3000, doc_base=1500, vocab_terms=2000, doc_len=24,
                                  passage_len=32), device="cpu", passages="global")
    assert torch.equal(half.vectors, full.vectors[1500:]) and torch.equal(half.passage_tok, full.passage_tok)
    assert int(full.bm25.off[-1]) == full.bm25.doc.numel()
    # postings are sorted inside every term even when built from several chunks
    off, doc = full.bm25.off.numpy(), full.bm25.doc.numpy()
    assert all((np.diff(doc[off[t]:off[t + 1]]) > 0).all() for t in range(0, 2000, 37))
    qt, qtok, qlen, dids = make_queries(full.cfg, 6, device="cpu")
    csr = dict(off=off, doc=doc, tf=full.bm25.tf.numpy(), doc_len=np.full(3000, 24, np.int32),
               df=full.df_local.numpy())
    for i in range(6):
        assert int(dids[i]) in [d for _, d in BM.bm25_ref(csr, qt[i].tolist(), k=3000)]


def test_fusion_oracles():
    a = torch.tensor([[1, 2, 3, -1]]); b = torch.tensor([[3, 9, -1, -1]])
    s, i = F.rrf_fuse_ref(a, b, 5)
    assert i[0].tolist() == [3, 1, 2, 9, -1] and s[0, 0] == pytest.approx(1 / 63 + 1 / 61)
    tok = torch.arange(40, dtype=torch.int32).reshape(4, 10); ln = torch.tensor([10, 3, 0, 10], dtype=torch.int32)
    ids, lens = F.build_pairs_ref(torch.tensor([[7, 8, 9]], dtype=torch.int32), torch.tensor([2], dtype=torch.int32),
                                  torch.tensor([[1, -1]]), tok, ln, 10, 12)
    assert ids[0].tolist() == [0, 7, 8, 2, 2, 10, 11, 12, 2, 1, 1, 1] and lens.tolist() == [9, 1]


def test_graft_build_entry():
    import __graft_entry__ as g

    g.build()
    assert _native.lib_path().exists()


_ASAN_SCENARIO = r"""
import ctypes, hashlib, random
import numpy as np
from infomesh_b200 import _native
from infomesh_b200.ops import bm25 as BM, dedup as DD

L = _native.lib()
L.im_md5_first8_be.restype = ctypes.c_ulonglong
for n in (0, 1, 55, 56, 57, 63, 64, 65, 119, 120, 1000, 4097):             # every MD5 padding boundary
    s = bytes(random.Random(n).randrange(256) for _ in range(n))
    assert L.im_md5_first8_be(s, ctypes.c_longlong(n)) == int.from_bytes(hashlib.md5(s).digest()[:8], "big")

b = BM.HostIndexBuilder()
rng = random.Random(7)
words = ["tensor", "memory", "Ünïcödé", "한국어", "x" * 300, "a", "", "tab\tsep", "new\nline", "\x00nul", "emoji😀", "mixed-CASE_42"]
for i in range(400):
    b.add_text(" ".join(rng.choice(words) for _ in range(rng.randrange(0, 40))))
b.add_text(""), b.add_text(" \t\n "), b.add_text("word " * 20000)
b.add_terms([0, 1, 2, 2, 2]), b.add_terms([])
assert b.lookup("TENSOR") >= 0 and b.lookup("") == -1 and b.lookup("z" * 5000) == -1
assert len(b.tokenize("tensor unknownword memory", add=False)) >= 2 and len(b.tokenize("", add=True)) == 0
for t in range(0, 5):
    assert isinstance(b.term(t), str)
csr = b.export()
assert csr["off"][-1] == len(csr["doc"]) == len(csr["tf"]) and len(csr["doc_len"]) == b.n_docs
b.close()

texts = ["", "one", "one two", "one two three", "The quick brown fox " * 500, "Ünïcödé wörds hérè", "a" * 10000, "한국어 문장 입니다"]
got = DD.simhash_cpu(texts)
assert (got == np.asarray([DD.simhash_py(t) for t in texts], dtype=np.uint64)).all()
table = np.asarray([rng.getrandbits(64) for _ in range(1000)] + [int(got[4])], dtype=np.uint64)
L.im_hamming_find_cpu.restype = ctypes.c_longlong
hit = L.im_hamming_find_cpu(table.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(len(table)), ctypes.c_ulonglong(int(got[4]) ^ 0b11), ctypes.c_int(3))
assert hit == 1000, hit
assert L.im_hamming_find_cpu(table.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(0), ctypes.c_ulonglong(1), ctypes.c_int(3)) == -1
print("ASAN_SCENARIO_OK")
"""


def test_host_runtime_is_clean_under_address_and_ub_sanitizers(tmp_path):
    """the C++ host library (tokeniser, index builder, MD5, SimHash, Hamming scan) built with
    -fsanitize=address,undefined and driven through the normal Python wrappers in a subprocess."""
    import os
    import subprocess
    import sys

    from infomesh_b200 import build as B

    runtime = B.sanitizer_runtime()
    if runtime is None:
        pytest.skip("no compiler with the sanitizer runtimes on this machine")
    try:
        lib = B.build_host_sanitized()
    except RuntimeError as exc:
        pytest.skip(f"sanitizer build unavailable: {exc}")
    env = dict(os.environ, INFOMESH_B200_NATIVE_LIB=str(lib), LD_PRELOAD=runtime, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0:exitcode=66",
               UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1", CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", _ASAN_SCENARIO], capture_output=True, text=True, env=env, timeout=600,
                         cwd=str(B.ROOT.parent))
    report = out.stdout[-2000:] + out.stderr[-6000:]
    assert out.returncode == 0 and "ASAN_SCENARIO_OK" in out.stdout, report
    assert "AddressSanitizer" not in out.stderr and "runtime error:" not in out.stderr, report
