"""LLM runtime clients against a local stand-in HTTP server, the summarisation engine, and explain mode
(model: reference tests/test_summarizer.py, test_llm_backends.py, test_explain.py)."""
import asyncio
import json
import threading
from http.server import BaseHTTPRequestHandler, HTTPServer

import pytest

from infomesh_b200.hashing import content_hash
from infomesh_b200.summarizer import engine as E


class _Runtime(BaseHTTPRequestHandler):
    """Speaks just enough of the Ollama, llama.cpp and vLLM HTTP dialects."""
    seen: list = []

    def log_message(self, *a):
        pass

    def _send(self, code, obj):
        raw = json.dumps(obj).encode()
        self.send_response(code)
        self.send_header("Content-Type", "application/json")
        self.send_header("Content-Length", str(len(raw)))
        self.end_headers()
        self.wfile.write(raw)

    def do_GET(self):
        routes = {"/api/tags": {"models": [{"name": "qwen2.5:3b"}, {"name": "llama3:8b"}]}, "/health": {"status": "ok"},
                  "/v1/models": {"data": [{"id": "qwen2.5-3b-instruct"}]}}
        if self.path in routes:
            self._send(200, routes[self.path])
        else:
            self._send(404, {})

    def do_POST(self):
        body = json.loads(self.rfile.read(int(self.headers.get("Content-Length", 0))) or b"{}")
        type(self).seen.append((self.path, body, self.headers.get("Authorization")))
        if self.path == "/api/generate":
            self._send(200, {"response": f"ollama:{body['options']['num_predict']}"})
        elif self.path == "/api/show":
            self._send(200 if body["name"].startswith("qwen") else 404, {"details": {"parameter_size": "3B", "quantization_level": "Q4_K_M"}})
        elif self.path == "/completion":
            self._send(200, {"content": " llama says hi \n"})
        elif self.path == "/v1/completions":
            self._send(200, {"choices": [{"text": "vllm text"}] if body["prompt"] != "empty" else []})
        else:
            self._send(500, {})


@pytest.fixture(scope="module")
def runtime_url():
    srv = HTTPServer(("127.0.0.1", 0), _Runtime)
    th = threading.Thread(target=srv.serve_forever, daemon=True)
    th.start()
    yield f"http://127.0.0.1:{srv.server_port}"
    srv.shutdown()
    th.join(timeout=5)


def test_ollama_client(runtime_url):
    b = E.OllamaBackend("qwen2.5:3b", runtime_url + "/")
    run = asyncio.run
    assert run(b.generate("hello", max_tokens=77)) == "ollama:77"
    path, body, _ = _Runtime.seen[-1]
    assert path == "/api/generate" and body["stream"] is False and body["options"]["temperature"] == 0.3
    assert run(b.is_available()) and not run(E.OllamaBackend("mistral:7b", runtime_url).is_available())
    info = run(b.model_info())
    assert (info.parameter_count, info.quantization, info.available, info.runtime) == ("3B", "Q4_K_M", True, E.LLMRuntime.OLLAMA)
    assert run(E.OllamaBackend("mistral:7b", runtime_url).model_info()).available is False


def test_llamacpp_and_vllm_clients(runtime_url):
    run = asyncio.run
    lc = E.LlamaCppBackend(runtime_url, "gguf-model")
    assert run(lc.generate("p")) == " llama says hi \n" and run(lc.is_available())
    assert _Runtime.seen[-1][1]["stop"] == ["\n\n---", "###"] and run(lc.model_info()).name == "gguf-model"
    v = E.VLLMBackend("qwen2.5:3b", runtime_url, api_key="sekrit")
    assert run(v.generate("p", max_tokens=9)) == "vllm text" and _Runtime.seen[-1][2] == "Bearer sekrit" and _Runtime.seen[-1][1]["max_tokens"] == 9
    assert run(v.generate("empty")) == "" and run(v.is_available()) and run(v.model_info()).runtime == E.LLMRuntime.VLLM
    assert not run(E.VLLMBackend("phi", runtime_url).is_available())


def test_http_errors_and_dead_runtimes(runtime_url):
    run = asyncio.run
    with pytest.raises(RuntimeError, match="HTTP 500"):
        run(E.LlamaCppBackend(runtime_url)._post_ok(runtime_url + "/nope", {}))
    dead = "http://127.0.0.1:9"
    assert not run(E.OllamaBackend(base_url=dead).is_available()) and not run(E.LlamaCppBackend(dead).is_available())
    assert not run(E.VLLMBackend(base_url=dead).is_available()) and run(E.OllamaBackend(base_url=dead).model_info()).available is False


def test_backend_factory_maps_runtime_names():
    assert isinstance(E.create_backend("ollama", "m"), E.OllamaBackend)
    for alias in ("llama.cpp", "llama_cpp", "llamacpp"):
        assert isinstance(E.create_backend(alias, "m", base_url="http://h:1/"), E.LlamaCppBackend)
    assert isinstance(E.create_backend("vllm"), E.VLLMBackend)
    native = E.create_backend("b200", "qwen2.5:3b")                     # an Ollama-style tag is not a T5 checkpoint
    assert isinstance(native, E.B200Backend) and native._model_name == "t5-small"
    assert E.create_backend("native", "t5-base")._model_name == "t5-base"
    with pytest.raises(ValueError, match="Unsupported LLM runtime"):
        E.create_backend("openai")
    info = asyncio.run(native.model_info())
    assert info.parameter_count == "60M" and info.quantization == "bf16" and info.runtime == E.LLMRuntime.B200


class Echo(E.LLMBackend):
    def __init__(self):
        self.prompts, self.infos = [], 0

    async def generate(self, prompt, *, max_tokens=512):
        self.prompts.append((prompt, max_tokens))
        return "  A short summary.  "

    async def is_available(self):
        return True

    async def model_info(self):
        self.infos += 1
        return E.ModelInfo("echo-1", E.LLMRuntime.OLLAMA, None, None, True)


def test_engine_builds_the_prompt_truncates_and_caches_model_info():
    be = Echo()
    eng = E.SummarizationEngine(be)
    text = "Q" * 9000
    r1 = asyncio.run(eng.summarize("https://e.org/a", "Title A", text, max_tokens=64))
    r2 = asyncio.run(eng.summarize("https://e.org/b", "Title B", "short body", max_input_chars=5))
    assert r1.summary == "A short summary." and r1.model == "echo-1" and r1.content_hash == content_hash(text) and r1.token_count == 5
    p1, mt = be.prompts[0]
    assert mt == 64 and "URL: https://e.org/a" in p1 and "Title: Title A" in p1 and p1.count("Q") == 8000 and "AI-generated" in p1
    assert "short" in be.prompts[1][0] and "short body" not in be.prompts[1][0]
    assert be.infos == 1 and eng.backend is be and asyncio.run(eng.is_available())
    assert E._estimate_tokens("") == 1 and r2.url == "https://e.org/b"


def test_native_backend_gets_a_bare_document_prompt():
    class Native(E.B200Backend):
        def __init__(self):
            super().__init__("t5-small")
            self.prompt = None

        async def generate(self, prompt, *, max_tokens=512):
            self.prompt = prompt
            return "s"

        async def is_available(self):
            return False

    be = Native()
    asyncio.run(E.SummarizationEngine(be).summarize("https://e.org/", "Kernel notes", "body text"))
    assert be.prompt == "Kernel notes\nbody text"


# ------------------------------------------------------------------ explain mode
def _rr(**kw):
    from infomesh_b200.index.ranking import RankedResult

    base = dict(doc_id=1, url="https://e.org/a", title="A", snippet="s", bm25_score=0.9, freshness_score=0.1, trust_score=0.85, authority_score=0.6,
                combined_score=0.7, crawled_at=0.0, title_match_score=0.7, url_path_score=0.4)
    base.update(kw)
    return RankedResult(**base)


def test_explain_breaks_a_score_into_weighted_signals():
    from infomesh_b200.index import ranking as R
    from infomesh_b200.search import explain as X

    e = X.explain_result(_rr())
    assert e.weights["bm25"] == R.WEIGHT_BM25 and e.weighted["bm25"] == pytest.approx(0.9 * R.WEIGHT_BM25)
    assert set(e.components) == {"bm25", "freshness", "trust", "authority", "title_match", "url_path"}
    assert e.notes == ["Strong keyword match", "Stale content — may need recrawl", "High-trust peer", "High domain authority", "Query matches title",
                       "Query matches URL path"]
    quiet = X.explain_result(_rr(bm25_score=0.5, freshness_score=0.5, trust_score=0.5, authority_score=0.1, title_match_score=0.0, url_path_score=0.0))
    assert quiet.notes == []
    assert X.explain_result(_rr(freshness_score=0.95)).notes[1] == "Recently crawled"
    d = e.to_dict()
    assert d["weighted_contributions"]["trust"] == round(0.85 * R.WEIGHT_TRUST, 4) and d["combined_score"] == 0.7


def test_explain_query_reports_the_pipeline_that_ran():
    from infomesh_b200.search import explain as X

    q = X.explain_query("Rust  ownership!", "rust ownership", [_rr(), _rr(url="https://e.org/b")], 12.345)
    assert q.pipeline == X.DEFAULT_PIPELINE and q.pipeline is not X.DEFAULT_PIPELINE and q.total_results == 2
    g = X.explain_query("q", "q", [], 1.0, pipeline=X.GPU_PIPELINE).to_dict()
    assert g["pipeline"][2] == "sim_topk_dense" and g["pipeline"][-1] == "rerank_select" and g["results"] == [] and g["elapsed_ms"] == 1.0
    assert q.to_dict()["elapsed_ms"] == 12.3 and len(q.to_dict()["results"]) == 2
