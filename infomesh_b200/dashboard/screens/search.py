"""Search tab: a query box over the local index; ``/`` focuses it, ``Esc`` hands the keyboard back to the tabs
(reference infomesh/dashboard/screens/search.py:22-216)."""
from __future__ import annotations

import contextlib

from textual.app import ComposeResult
from textual.containers import Vertical
from textual.widgets import Input, Static


def _markup(text: str) -> str:
    return text.replace("[", "\\[")


class SearchResultsPanel(Static):
    """Ranked results with score and highlighted snippet."""

    def __init__(self, config=None, **kw):
        super().__init__("[dim]Type a query above and press Enter to search the local index[/]", **kw)
        self.config = config

    @staticmethod
    def render_results(query: str, results: list[dict[str, object]], elapsed_ms: float, source: str = "local") -> str:
        if not results:
            return "No results found."
        rows = [f"[dim]{len(results)} results for \"{_markup(query)}\" in {elapsed_ms:.0f} ms ({source})[/]"]
        for i, r in enumerate(results, 1):
            snip = _markup(str(r.get("snippet", ""))).replace("<b>", "[bold yellow]").replace("</b>", "[/]").replace("\n", " ")[:180]
            rows.append(f"[bold cyan]{i}. {_markup(str(r.get('title') or r.get('url', '')))}[/]  [dim]{float(r.get('score', 0.0) or 0.0):.3f}[/]\n"
                        f"   {r.get('url', '')}\n   {snip}")
        return "\n".join(rows)

    def on_mount(self) -> None:
        self.update("[dim]Type a query above and press Enter to search the local index[/]")

    def display_error(self, message: str) -> None:
        self.update(f"[bold red]Error:[/] {_markup(message)}")

    def display_results(self, query: str, results: list[dict[str, object]], elapsed_ms: float, source: str = "local") -> None:
        self.update(self.render_results(query, results, elapsed_ms, source))


class SearchPane(Vertical):
    BINDINGS = [("slash", "focus_search", "Search"), ("escape", "blur_search", "Back")]

    def __init__(self, config, **kw):
        super().__init__(**kw)
        self.config = config

    def compose(self) -> ComposeResult:
        yield Input(placeholder="search the local index and press Enter", id="se-input")
        yield SearchResultsPanel(self.config, id="se-results")

    def action_focus_search(self) -> None:
        with contextlib.suppress(Exception):
            self.query_one("#se-input", Input).focus()

    def action_blur_search(self) -> None:
        with contextlib.suppress(Exception):
            self.screen.focus_next()

    def refresh_data(self) -> None:
        """Nothing to poll: this tab only changes when the user submits a query."""

    def run_query(self, query: str) -> str:
        from infomesh_b200.index.local_store import LocalStore
        from infomesh_b200.search.query import search_local

        cfg = self.config
        with LocalStore(db_path=cfg.index.db_path, compression_enabled=cfg.storage.compression_enabled,
                        compression_level=cfg.storage.compression_level) as st:
            res = search_local(st, query, limit=8)
        rows = [{"title": r.title, "url": r.url, "snippet": r.snippet, "score": r.combined_score} for r in res.results]
        return SearchResultsPanel.render_results(query, rows, res.elapsed_ms)

    def on_input_submitted(self, event: Input.Submitted) -> None:
        q = event.value.strip()
        if q:
            panel = self.query_one(SearchResultsPanel)
            try:
                panel.update(self.run_query(q))
            except Exception as exc:  # noqa: BLE001
                panel.display_error(f"search failed: {exc}")
