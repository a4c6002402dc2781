"""Network tab: P2P state, peers, DHT and bandwidth counters (reference infomesh/dashboard/screens/network.py:26-382)."""
from __future__ import annotations

from textual.app import ComposeResult
from textual.containers import Vertical
from textual.widgets import Static

from infomesh_b200.dashboard import utils as U
from infomesh_b200.dashboard.widgets import SparklineChart


class NetworkPane(Vertical):
    def __init__(self, config, **kw):
        super().__init__(**kw)
        self.config = config
        self._last = (0, 0)

    def compose(self) -> ComposeResult:
        yield Static("", id="nw-state")
        yield SparklineChart("upload", color="yellow", id="nw-up")
        yield SparklineChart("download", color="green", id="nw-down")
        yield Static("", id="nw-peers")

    def on_mount(self) -> None:
        self.refresh_data()
        self.set_interval(2.0, self.refresh_data)

    def refresh_data(self) -> None:
        st, cfg = U.read_p2p_status(self.config), self.config
        state = str(st.get("state", "stopped")) if st else "not started"
        dht, bw, boot = (st.get(k) if isinstance(st.get(k), dict) else {} for k in ("dht", "bandwidth", "bootstrap")) if st else ({}, {}, {})
        lines = [f"P2P [bold]{state}[/]  ·  peers [bold]{int(st.get('peers', 0) or 0) if st else 0}[/]  ·  port {cfg.node.listen_port}/tcp  ·  "
                 f"replication {cfg.network.replication_factor}x",
                 f"DHT  stored {dht.get('keys_stored', 0):,}  published {dht.get('keys_published', 0):,}  gets {dht.get('gets_performed', 0):,}  "
                 f"puts {dht.get('puts_performed', 0):,}",
                 f"bootstrap  {boot.get('connected', 0)} connected of {boot.get('configured', len(cfg.network.bootstrap_nodes))} configured",
                 "listen  " + ", ".join(str(a) for a in (st.get("listen_addrs", []) if st else [])[:2])]
        self.query_one("#nw-state", Static).update("\n".join(lines))
        up, down = int(bw.get("upload_bytes", 0) or 0), int(bw.get("download_bytes", 0) or 0)
        self.query_one("#nw-up", SparklineChart).push(max(0, up - self._last[0]) / 2048)
        self.query_one("#nw-down", SparklineChart).push(max(0, down - self._last[1]) / 2048)
        self._last = (up, down)
        ids = st.get("peer_ids", []) if st else []
        vers = st.get("peer_versions", {}) if st else {}
        self.query_one("#nw-peers", Static).update("[bold]Peers[/]\n" + ("\n".join(f"  {p[:24]}…  v{vers.get(p, '?')}" for p in ids[:15]) or "  [dim]none connected[/]"))
