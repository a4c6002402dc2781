"""Model-side tokenisers.

There is no network access for vocabulary files, so the default is a deterministic **hash word-piece**
tokeniser: words (``\\w+`` runs, lower-cased) map to ``first_id + fnv1a(word) % (vocab - first_id)``.
It has the right shape (one id per word, special tokens, truncation, padding) for the randomly initialised
models; a real ``vocab.txt`` can be supplied to use exact WordPiece lookup instead.
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from pathlib import Path

_WORD = re.compile(r"\w+", re.UNICODE)


def fnv1a(data: bytes) -> int:
    h = 0xCBF29CE484222325
    for b in data:
        h ^= b
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


@dataclass
class SpecialTokens:
    cls: int = 101
    sep: int = 102
    pad: int = 0
    unk: int = 100
    first_regular: int = 1000


BERT_SPECIALS = SpecialTokens()
XLMR_SPECIALS = SpecialTokens(cls=0, sep=2, pad=1, unk=3, first_regular=1000)
T5_SPECIALS = SpecialTokens(cls=0, sep=1, pad=0, unk=2, first_regular=1000)


class HashTokenizer:
    def __init__(self, vocab_size: int, specials: SpecialTokens = BERT_SPECIALS, vocab_file: str | None = None):
        self.vocab_size = vocab_size
        self.sp = specials
        self.vocab: dict[str, int] | None = None
        if vocab_file and Path(vocab_file).exists():
            self.vocab = {w.rstrip("\n"): i for i, w in enumerate(Path(vocab_file).read_text("utf-8").splitlines())}

    def word_id(self, word: str) -> int:
        if self.vocab is not None:
            return self.vocab.get(word, self.sp.unk)
        span = self.vocab_size - self.sp.first_regular
        return self.sp.first_regular + fnv1a(word.encode("utf-8")) % span

    def words(self, text: str) -> list[str]:
        return _WORD.findall(text.lower())

    def encode(self, text: str, max_len: int = 512, add_special: bool = True) -> list[int]:
        ids = [self.word_id(w) for w in self.words(text)]
        if add_special:
            ids = [self.sp.cls] + ids[:max_len - 2] + [self.sp.sep]
        return ids[:max_len]

    def encode_plain(self, text: str, max_len: int) -> list[int]:
        return [self.word_id(w) for w in self.words(text)][:max_len]

    def encode_batch(self, texts: list[str], max_len: int = 512, pad_to_multiple: int = 8):
        """-> (ids int32 [B, S], lengths int32 [B]) as torch CPU tensors (pinned when CUDA is present)."""
        import torch

        enc = [self.encode(t, max_len) for t in texts]
        longest = max((len(e) for e in enc), default=1)
        S = min(max_len, ((longest + pad_to_multiple - 1) // pad_to_multiple) * pad_to_multiple)
        ids = torch.full((len(enc), S), self.sp.pad, dtype=torch.int32)
        lens = torch.zeros((len(enc),), dtype=torch.int32)
        for i, e in enumerate(enc):
            ids[i, :len(e)] = torch.tensor(e, dtype=torch.int32)
            lens[i] = len(e)
        if torch.cuda.is_available():
            ids, lens = ids.pin_memory(), lens.pin_memory()
        return ids, lens

    def decode(self, ids) -> str:
        if self.vocab is None:
            return " ".join(f"<{int(i)}>" for i in ids)
        inv = getattr(self, "_inv", None)
        if inv is None:
            inv = self._inv = {i: w for w, i in self.vocab.items()}
        return " ".join(inv.get(int(i), "[UNK]") for i in ids)


class Seq2SeqTokenizer:
    """T5-side tokeniser: SentencePiece when ``spiece.model`` is available, hash word ids otherwise.
    ``encode`` appends EOS (id 1); ``decode`` stops at EOS and drops padding."""

    def __init__(self, vocab_size: int = 32128, model_file: str | None = None):
        self.vocab_size = vocab_size
        self._sp = None
        self._hash = HashTokenizer(vocab_size, T5_SPECIALS)
        self._seen: dict[int, str] = {}
        if model_file and Path(model_file).exists():
            try:
                import sentencepiece as spm

                self._sp = spm.SentencePieceProcessor(model_file=str(model_file))
            except Exception:  # noqa: BLE001 — fall back to hash ids
                self._sp = None

    def encode(self, text: str, max_len: int = 512, add_eos: bool = True) -> list[int]:
        if self._sp is not None:
            ids = list(self._sp.encode(text))
        else:
            words = self._hash.words(text)
            ids = [self._hash.word_id(w) for w in words]
            for i, w in zip(ids, words):
                self._seen.setdefault(i, w)       # lets decode() echo words this process has encoded
        ids = ids[:max_len - 1 if add_eos else max_len]
        return ids + [1] if add_eos else ids

    def decode(self, ids) -> str:
        toks = []
        for i in ids:
            i = int(i)
            if i == 1:
                break
            if i != 0:
                toks.append(i)
        if self._sp is not None:
            return self._sp.decode(toks)
        return " ".join(self._seen.get(i, f"<{i}>") for i in toks)


def load_tokenizer(name_or_dir: str, vocab_size: int = 32128) -> Seq2SeqTokenizer:
    p = Path(name_or_dir)
    model_file = str(p / "spiece.model") if p.is_dir() else None
    return Seq2SeqTokenizer(vocab_size, model_file)


class WordPieceTokenizer(HashTokenizer):
    """BERT WordPiece from a ``vocab.txt``: lower-case, split on whitespace and punctuation, then greedy
    longest-match-first with ``##`` continuation pieces; words that cannot be segmented map to ``[UNK]``.
    Same interface as :class:`HashTokenizer` (``encode`` / ``encode_plain`` / ``encode_batch``)."""

    _SPLIT = re.compile(r"\w+|[^\w\s]", re.UNICODE)

    def __init__(self, vocab_file, max_chars_per_word: int = 100):
        lines = Path(vocab_file).read_text("utf-8").splitlines()
        self.vocab = {w: i for i, w in enumerate(lines)}
        self.vocab_size = len(lines)
        g = self.vocab.get
        self.sp = SpecialTokens(cls=g("[CLS]", 101), sep=g("[SEP]", 102), pad=g("[PAD]", 0), unk=g("[UNK]", 100), first_regular=0)
        self.max_chars = max_chars_per_word

    def pieces(self, word: str) -> list[int]:
        if len(word) > self.max_chars:
            return [self.sp.unk]
        out, start = [], 0
        while start < len(word):
            end, cur = len(word), None
            while start < end:
                sub = word[start:end] if start == 0 else "##" + word[start:end]
                if sub in self.vocab:
                    cur = self.vocab[sub]
                    break
                end -= 1
            if cur is None:
                return [self.sp.unk]
            out.append(cur)
            start = end
        return out

    def _ids(self, text: str) -> list[int]:
        ids: list[int] = []
        for w in self._SPLIT.findall(text.lower()):
            ids.extend(self.pieces(w))
        return ids

    def word_id(self, word: str) -> int:
        p = self.pieces(word.lower())
        return p[0] if p else self.sp.unk

    def encode(self, text: str, max_len: int = 512, add_special: bool = True) -> list[int]:
        ids = self._ids(text)
        if add_special:
            ids = [self.sp.cls] + ids[:max_len - 2] + [self.sp.sep]
        return ids[:max_len]

    def encode_plain(self, text: str, max_len: int) -> list[int]:
        return self._ids(text)[:max_len]


class SentencePieceTokenizer(HashTokenizer):
    """XLM-R SentencePiece (``sentencepiece.bpe.model``) with fairseq's id shift (<s>=0, <pad>=1, </s>=2, <unk>=3, piece
    ids + 1)."""

    def __init__(self, model_file, vocab_size: int):
        import sentencepiece as spm

        self.proc = spm.SentencePieceProcessor(model_file=str(model_file))
        self.vocab_size = vocab_size
        self.vocab = None
        self.sp = XLMR_SPECIALS

    def _ids(self, text: str) -> list[int]:
        return [i + 1 if i > 0 else self.sp.unk for i in self.proc.encode(text)]

    def word_id(self, word: str) -> int:
        ids = self._ids(word)
        return ids[0] if ids else self.sp.unk

    def encode(self, text: str, max_len: int = 512, add_special: bool = True) -> list[int]:
        ids = self._ids(text)
        if add_special:
            ids = [self.sp.cls] + ids[:max_len - 2] + [self.sp.sep]
        return ids[:max_len]

    def encode_plain(self, text: str, max_len: int) -> list[int]:
        return self._ids(text)[:max_len]
