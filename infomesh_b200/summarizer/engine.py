"""Summarisation engine with pluggable backends (reference infomesh/summarizer/engine.py:27-436).

Backends: ``ollama`` (``/api/generate``), ``llama.cpp`` (``/completion``), ``vllm`` (OpenAI ``/v1/completions``) —
the three out-of-process runtimes of the reference, spoken to with the standard library so no HTTP client package is
needed — plus ``b200``: the in-process T5 encoder-decoder running on this repo's own kernels
(:mod:`infomesh_b200.models.t5`), which is what replaces the external GPU servers on a B200 node.
"""
from __future__ import annotations

import abc
import asyncio
import json
import time
import urllib.error
import urllib.request
from dataclasses import dataclass
from enum import StrEnum
from typing import Any

from infomesh_b200.hashing import content_hash
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


class LLMRuntime(StrEnum):
    OLLAMA = "ollama"
    LLAMA_CPP = "llama.cpp"
    VLLM = "vllm"
    B200 = "b200"


@dataclass(frozen=True)
class SummaryResult:
    url: str
    summary: str
    model: str
    runtime: LLMRuntime
    content_hash: str
    elapsed_ms: float
    token_count: int | None


@dataclass(frozen=True)
class ModelInfo:
    name: str
    runtime: LLMRuntime
    parameter_count: str | None
    quantization: str | None
    available: bool


class LLMBackend(abc.ABC):
    @abc.abstractmethod
    async def generate(self, prompt: str, *, max_tokens: int = 512) -> str: ...

    @abc.abstractmethod
    async def is_available(self) -> bool: ...

    @abc.abstractmethod
    async def model_info(self) -> ModelInfo: ...

    async def close(self) -> None:
        return None


class _HttpJsonBackend(LLMBackend):
    """Blocking urllib calls moved off the loop with ``to_thread``; 120 s timeout like the reference client."""
    timeout = 120.0

    def _call(self, method: str, url: str, body: dict | None = None, headers: dict[str, str] | None = None) -> tuple[int, Any]:
        data = json.dumps(body).encode() if body is not None else None
        hdr = {"Content-Type": "application/json", **(headers or {})}
        req = urllib.request.Request(url, data=data, headers=hdr, method=method)
        try:
            with urllib.request.urlopen(req, timeout=self.timeout) as resp:  # noqa: S310 — operator-configured local runtime
                raw = resp.read()
                return resp.status, (json.loads(raw) if raw else {})
        except urllib.error.HTTPError as exc:
            return exc.code, {}

    async def _request(self, method: str, url: str, body: dict | None = None, headers: dict[str, str] | None = None):
        return await asyncio.to_thread(self._call, method, url, body, headers)

    async def _post_ok(self, url: str, body: dict, headers: dict[str, str] | None = None) -> Any:
        status, data = await self._request("POST", url, body, headers)
        if status >= 400:
            raise RuntimeError(f"LLM runtime returned HTTP {status} for {url}")
        return data


class OllamaBackend(_HttpJsonBackend):
    def __init__(self, model: str = "qwen2.5:3b", base_url: str = "http://localhost:11434"):
        self._model, self._base_url = model, base_url.rstrip("/")

    async def generate(self, prompt: str, *, max_tokens: int = 512) -> str:
        data = await self._post_ok(f"{self._base_url}/api/generate",
                                   {"model": self._model, "prompt": prompt, "stream": False,
                                    "options": {"num_predict": max_tokens, "temperature": 0.3}})
        return str(data["response"])

    async def is_available(self) -> bool:
        try:
            status, data = await self._request("GET", f"{self._base_url}/api/tags")
        except (OSError, ValueError):
            return False
        family = self._model.split(":")[0]
        return status == 200 and any(str(m.get("name", "")).startswith(family) for m in data.get("models", []))

    async def model_info(self) -> ModelInfo:
        try:
            status, info = await self._request("POST", f"{self._base_url}/api/show", {"name": self._model})
            if status == 200:
                d = info.get("details", {})
                return ModelInfo(self._model, LLMRuntime.OLLAMA, d.get("parameter_size"), d.get("quantization_level"), True)
        except (OSError, ValueError):
            pass
        return ModelInfo(self._model, LLMRuntime.OLLAMA, None, None, False)


class LlamaCppBackend(_HttpJsonBackend):
    def __init__(self, base_url: str = "http://localhost:8080", model_name: str = "llama-cpp-model"):
        self._base_url, self._model_name = base_url.rstrip("/"), model_name

    async def generate(self, prompt: str, *, max_tokens: int = 512) -> str:
        data = await self._post_ok(f"{self._base_url}/completion",
                                   {"prompt": prompt, "n_predict": max_tokens, "temperature": 0.3, "stop": ["\n\n---", "###"]})
        return str(data["content"])

    async def is_available(self) -> bool:
        try:
            return (await self._request("GET", f"{self._base_url}/health"))[0] == 200
        except (OSError, ValueError):
            return False

    async def model_info(self) -> ModelInfo:
        return ModelInfo(self._model_name, LLMRuntime.LLAMA_CPP, None, None, await self.is_available())


class VLLMBackend(_HttpJsonBackend):
    def __init__(self, model: str = "qwen2.5:3b", base_url: str = "http://localhost:8000", *, api_key: str = "EMPTY"):
        self._model, self._base_url, self._api_key = model, base_url.rstrip("/"), api_key

    async def generate(self, prompt: str, *, max_tokens: int = 512) -> str:
        data = await self._post_ok(f"{self._base_url}/v1/completions",
                                   {"model": self._model, "prompt": prompt, "max_tokens": max_tokens, "temperature": 0.3},
                                   {"Authorization": f"Bearer {self._api_key}"})
        choices = data.get("choices", [])
        return str(choices[0].get("text", "")) if choices else ""

    async def is_available(self) -> bool:
        try:
            status, data = await self._request("GET", f"{self._base_url}/v1/models")
        except (OSError, ValueError):
            return False
        family = self._model.split(":")[0]
        return status == 200 and any(str(m.get("id", "")).startswith(family) for m in data.get("data", []))

    async def model_info(self) -> ModelInfo:
        return ModelInfo(self._model, LLMRuntime.VLLM, None, None, await self.is_available())


class B200Backend(LLMBackend):
    """In-process T5 summariser on the native kernels.  ``model`` is a T5 config name (``t5-small``) or a directory
    with ``config.json`` + weight shards; without weights the model is random-initialised (benchmarks / tests)."""

    def __init__(self, model: str = "t5-small", *, device: str | None = None, max_input_tokens: int = 512):
        self._model_name, self._device, self._max_in = model, device, max_input_tokens
        self._model = None
        self._tok = None
        self._lock = asyncio.Lock()

    def _load(self):
        if self._model is None:
            import torch

            from infomesh_b200.models.t5 import T5Model, load_t5
            from infomesh_b200.utils.tokenizer import load_tokenizer

            dev = torch.device(self._device or "cuda:0")
            self._model = load_t5(self._model_name, device=dev)
            self._tok = load_tokenizer(self._model_name, vocab_size=self._model.cfg.vocab_size)
            assert isinstance(self._model, T5Model)
        return self._model, self._tok

    def _generate_sync(self, prompt: str, max_tokens: int) -> str:
        import torch

        model, tok = self._load()
        ids = tok.encode("summarize: " + prompt, max_len=self._max_in, add_eos=True)
        inp = torch.tensor([ids], dtype=torch.int32, device=model.device)
        lens = torch.tensor([len(ids)], dtype=torch.int32, device=model.device)
        out = model.generate(inp, lens, max_new_tokens=min(max_tokens, 256))
        return tok.decode(out[0].tolist())

    async def generate(self, prompt: str, *, max_tokens: int = 512) -> str:
        async with self._lock:            # one decode stream per device
            return await asyncio.to_thread(self._generate_sync, prompt, max_tokens)

    async def is_available(self) -> bool:
        try:
            import torch

            from infomesh_b200 import _native

            return bool(torch.cuda.is_available() and _native.available())
        except Exception:  # noqa: BLE001
            return False

    async def model_info(self) -> ModelInfo:
        return ModelInfo(self._model_name, LLMRuntime.B200, "60M" if "small" in self._model_name else None, "bf16",
                         await self.is_available())


_SUMMARIZE_PROMPT = """\
You are a precise summarization assistant. Summarize the following web page content.
Keep the summary concise (3-5 sentences), factual, and information-dense.
Do not add opinions or information not present in the text.
Label this output as AI-generated.

URL: {url}
Title: {title}

Content:
{text}

Summary:"""


def create_backend(runtime: str, model: str = "qwen2.5:3b", *, base_url: str | None = None) -> LLMBackend:
    kw = {"base_url": base_url} if base_url is not None else {}
    match runtime:
        case "ollama":
            return OllamaBackend(model=model, **kw)
        case "llama.cpp" | "llama_cpp" | "llamacpp":
            return LlamaCppBackend(model_name=model, **kw)
        case "vllm":
            return VLLMBackend(model=model, **kw)
        case "b200" | "native":
            return B200Backend(model=model if model and ":" not in model else "t5-small")
        case _:
            raise ValueError(f"Unsupported LLM runtime: {runtime!r}. Use 'ollama', 'llama.cpp', 'vllm' or 'b200'.")


def _estimate_tokens(text: str) -> int:
    return max(1, len(text) // 4)


class SummarizationEngine:
    def __init__(self, backend: LLMBackend):
        self._backend = backend
        self._info: ModelInfo | None = None

    @property
    def backend(self) -> LLMBackend:
        return self._backend

    async def summarize(self, url: str, title: str, text: str, *, max_tokens: int = 512,
                        max_input_chars: int = 8000) -> SummaryResult:
        body = text[:max_input_chars]
        # the native seq2seq model is trained on "summarize: <document>", not on an instruction preamble
        prompt = (f"{title}\n{body}" if isinstance(self._backend, B200Backend)
                  else _SUMMARIZE_PROMPT.format(url=url, title=title, text=body))
        t0 = time.monotonic()
        summary = await self._backend.generate(prompt, max_tokens=max_tokens)
        ms = (time.monotonic() - t0) * 1000
        if self._info is None:
            self._info = await self._backend.model_info()
        res = SummaryResult(url, summary.strip(), self._info.name, self._info.runtime, content_hash(text), round(ms, 1),
                            _estimate_tokens(summary))
        logger.info("summarization_complete", url=url, model=res.model, elapsed_ms=res.elapsed_ms, summary_length=len(res.summary))
        return res

    async def is_available(self) -> bool:
        return await self._backend.is_available()
