"""Command line: ``start stop update status crawl mcp dashboard search index config keys peer feeds feedback doctor
bench`` (+ hidden ``_serve``) — the verb set of reference infomesh/cli/__init__.py:49-159, built from one module per
concern."""
from __future__ import annotations

import click

from infomesh_b200 import __version__


@click.group()
@click.version_option(version=__version__, prog_name="infomesh")
def cli() -> None:
    """InfoMesh — decentralized P2P search engine for LLMs via MCP (B200-native build)."""


from infomesh_b200.cli.config import config_group  # noqa: E402
from infomesh_b200.cli.crawl import crawl, dashboard, feeds_group, mcp_cmd  # noqa: E402
from infomesh_b200.cli.index import index_group  # noqa: E402
from infomesh_b200.cli.keys import keys_group  # noqa: E402
from infomesh_b200.cli.peer import peer_group  # noqa: E402
from infomesh_b200.cli.search import feedback_group, search  # noqa: E402
from infomesh_b200.cli.serve import serve, start, status, stop, update  # noqa: E402

for _cmd in (start, stop, update, serve, status, crawl, mcp_cmd, dashboard, search, index_group, config_group, keys_group, peer_group,
             feeds_group, feedback_group):
    cli.add_command(_cmd)


@cli.command()
def doctor() -> None:
    """Run diagnostic checks on the installation (data dir, keys, index, ports, disk, GPU)."""
    from infomesh_b200.config import load_config
    from infomesh_b200.diagnostics import run_diagnostics

    cfg = load_config()
    report = run_diagnostics(cfg.node.data_dir, p2p_port=cfg.node.listen_port)
    icon = {"ok": ("✔", "green"), "warning": ("⚠", "yellow"), "error": ("✖", "red")}
    click.echo("InfoMesh Doctor\n" + "=" * 40)
    for c in report.checks:
        mark, color = icon.get(c.status, ("?", None))
        click.secho(f"  {mark} {c.name}: {c.message}", fg=color)
    click.secho(f"\nSummary: {report.summary}", fg="green" if report.ok else "yellow", bold=True)


@cli.command()
@click.option("--iterations", "-n", default=50, help="Iterations per benchmark")
@click.option("--gpu", is_flag=True, help="Also time the device kernels (needs a B200)")
def bench(iterations: int, gpu: bool) -> None:
    """Run micro-benchmarks of the query-side text pipeline (and, with --gpu, of the device kernels)."""
    from infomesh_b200.benchmarks import BenchmarkSuite, benchmark
    from infomesh_b200.search.cjk import is_cjk_text, tokenize_query_cjk
    from infomesh_b200.search.nlp import expand_query, parse_natural_query
    from infomesh_b200.search.passage import split_passages
    from infomesh_b200.search.quality import QueryIntentClassifier

    suite = BenchmarkSuite()
    for name, fn, arg in (("query_expansion", expand_query, "python async error"),
                          ("nlp_parse", parse_natural_query, "python tutorial last week site:docs.python.org"),
                          ("passage_split", split_passages, "Hello world. " * 100), ("cjk_detect", is_cjk_text, "中文测试文本"),
                          ("cjk_tokenize", tokenize_query_cjk, "中文搜索测试"),
                          ("intent_classify", QueryIntentClassifier().classify, "how to install python")):
        suite.add(benchmark(fn, arg, iterations=iterations, name=name))
    if gpu:
        import torch

        from infomesh_b200.benchmarks import benchmark_cuda
        from infomesh_b200.ops.search import sim_topk

        q = torch.nn.functional.normalize(torch.randn(64, 384, device="cuda"), dim=1).bfloat16()
        d = torch.nn.functional.normalize(torch.randn(1_000_000, 384, device="cuda"), dim=1).bfloat16()
        suite.add(benchmark_cuda(lambda: sim_topk(q, d, 10), iterations=iterations, name="sim_topk_64x1M"))
    click.echo(suite.report())
