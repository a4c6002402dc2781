"""Stand-in for the `sentence-transformers` wheel (absent from this image) so the reference's UNMODIFIED
``VectorStore._get_embedder / _embed`` runs.

``SentenceTransformer(name).encode(texts)`` is a HuggingFace ``transformers.BertModel`` of the named architecture with
random-initialised weights (there is no network for checkpoints -- the repo's own arm is random-init too), run in fp32
by PyTorch's stock kernels on CUDA when available (what sentence-transformers does), pooled and L2-normalised.  With no
vocabulary files offline, words are hashed into the model's id space.  No kernel, model or engine of infomesh_b200 is
used."""
from __future__ import annotations

import re
import zlib

import numpy as np
import torch

_ARCH = {
    # name fragment -> (layers, hidden, heads, ffn, vocab, max_seq, pooling)
    "all-minilm-l6": (6, 384, 12, 1536, 30522, 256, "mean"),
    "bge-small": (12, 384, 12, 1536, 30522, 512, "cls"),
    "bge-base": (12, 768, 12, 3072, 30522, 512, "cls"),
}
_WORD = re.compile(r"\w+", re.UNICODE)


class SentenceTransformer:
    def __init__(self, model_name_or_path: str = "all-MiniLM-L6-v2", device: str | None = None, **_k):
        from transformers import BertConfig, BertModel

        key = next((k for k in _ARCH if k in model_name_or_path.lower()), "all-minilm-l6")
        L, H, A, F, V, S, pool = _ARCH[key]
        self.max_seq_length = S
        self._pool = pool
        self._vocab = V
        self.device = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
        torch.manual_seed(1234)
        cfg = BertConfig(vocab_size=V, hidden_size=H, num_hidden_layers=L, num_attention_heads=A, intermediate_size=F,
                         max_position_embeddings=512)
        self._model = BertModel(cfg, add_pooling_layer=False).eval().to(self.device)

    def get_sentence_embedding_dimension(self) -> int:
        return self._model.config.hidden_size

    def _tokenize(self, text: str) -> list[int]:
        ids = [101]
        for w in _WORD.findall(text.lower()):
            ids.append(1000 + zlib.crc32(w.encode("utf-8")) % (self._vocab - 1000))
            if len(ids) >= self.max_seq_length - 1:
                break
        ids.append(102)
        return ids

    @torch.no_grad()
    def encode(self, sentences, batch_size: int = 32, show_progress_bar: bool = False, convert_to_numpy: bool = True,
               normalize_embeddings: bool = False, **_k):
        single = isinstance(sentences, str)
        texts = [sentences] if single else list(sentences)
        outs = []
        for a in range(0, len(texts), batch_size):
            toks = [self._tokenize(t) for t in texts[a:a + batch_size]]
            S = max(len(t) for t in toks)
            ids = torch.zeros((len(toks), S), dtype=torch.long)
            mask = torch.zeros((len(toks), S), dtype=torch.long)
            for i, t in enumerate(toks):
                ids[i, :len(t)] = torch.tensor(t)
                mask[i, :len(t)] = 1
            ids, mask = ids.to(self.device), mask.to(self.device)
            h = self._model(input_ids=ids, attention_mask=mask).last_hidden_state
            if self._pool == "cls":
                e = h[:, 0]
            else:
                m = mask[..., None].to(h.dtype)
                e = (h * m).sum(1) / m.sum(1).clamp(min=1)
            e = torch.nn.functional.normalize(e, dim=1)
            outs.append(e.float().cpu())
        emb = torch.cat(outs).numpy() if outs else np.zeros((0, self.get_sentence_embedding_dimension()), np.float32)
        return emb[0] if single else emb
