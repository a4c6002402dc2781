"""Network tab: P2P state, DHT counters, bandwidth sparklines, peer table
(reference infomesh/dashboard/screens/network.py:26-382).  All four panels read the node's ``p2p_status.json``."""
from __future__ import annotations

from textual.app import ComposeResult
from textual.containers import Vertical
from textual.widgets import Static

from infomesh_b200.dashboard import utils as U
from infomesh_b200.dashboard.widgets import SparklineChart


def _section(st: dict, key: str) -> dict:
    v = st.get(key) if st else None
    return v if isinstance(v, dict) else {}


class P2PStatusPanel(Static):
    def __init__(self, config, **kw):
        super().__init__("", **kw)
        self.config = config

    def on_mount(self) -> None:
        self.show({})

    def update_status(self, data: dict[str, object], countdown: int = 0) -> None:
        self.show(dict(data or {}))

    def show(self, st: dict) -> None:
        cfg, boot = self.config, _section(st, "bootstrap")
        state = str(st.get("state", "stopped")) if st else "not started"
        lines = [f"P2P [bold]{state}[/]  ·  peers [bold]{int(st.get('peers', 0) or 0) if st else 0}[/]  ·  port {cfg.node.listen_port}/tcp  ·  "
                 f"replication {cfg.network.replication_factor}x",
                 f"bootstrap  {boot.get('connected', 0)} connected of {boot.get('configured', len(cfg.network.bootstrap_nodes))} configured",
                 "listen  " + ", ".join(str(a) for a in (st.get("listen_addrs", []) if st else [])[:2])]
        self.update("\n".join(lines))


class DHTPanel(Static):
    def on_mount(self) -> None:
        self.show({})

    def update_data(self, dht_data: dict[str, int], *, p2p_state: str = "stopped") -> None:
        self.show({"dht": dict(dht_data or {}), "state": p2p_state})

    def show(self, st: dict) -> None:
        dht = _section(st, "dht")
        self.update(f"DHT  stored {dht.get('keys_stored', 0):,}  published {dht.get('keys_published', 0):,}  gets {dht.get('gets_performed', 0):,}  "
                f"puts {dht.get('puts_performed', 0):,}")


class BandwidthPanel(Vertical):
    """Upload / download rate sparklines from the cumulative byte counters (KiB/s at the 2 s refresh)."""

    def __init__(self, **kw):
        super().__init__(**kw)
        self.last = (0, 0)

    def compose(self) -> ComposeResult:
        yield SparklineChart("upload", color="yellow", id="nw-up")
        yield SparklineChart("download", color="green", id="nw-down")

    def update_from_status(self, bw_data: dict[str, int]) -> None:
        self.show({"bandwidth": dict(bw_data or {})})

    def show(self, st: dict) -> None:
        bw = _section(st, "bandwidth")
        up, down = int(bw.get("upload_bytes", 0) or 0), int(bw.get("download_bytes", 0) or 0)
        self.query_one("#nw-up", SparklineChart).push(max(0, up - self.last[0]) / 2048)
        self.query_one("#nw-down", SparklineChart).push(max(0, down - self.last[1]) / 2048)
        self.last = (up, down)


class PeerTable(Static):
    def set_peers(self, peer_ids: list[str], versions: dict[str, str] | None = None) -> None:
        self.show({"peer_ids": list(peer_ids), "peer_versions": dict(versions or {})})

    def show(self, st: dict) -> None:
        ids = st.get("peer_ids", []) if st else []
        vers = st.get("peer_versions", {}) if st else {}
        self.update("[bold]Peers[/]\n" + ("\n".join(f"  {p[:24]}…  v{vers.get(p, '?')}" for p in ids[:15]) or "  [dim]none connected[/]"))


class NetworkPane(Vertical):
    def __init__(self, config, **kw):
        super().__init__(**kw)
        self.config = config

    @property
    def _last(self) -> tuple[int, int]:
        return self.query_one(BandwidthPanel).last

    def compose(self) -> ComposeResult:
        yield P2PStatusPanel(self.config, id="nw-state")
        yield DHTPanel("", id="nw-dht")
        yield BandwidthPanel(id="nw-bw")
        yield PeerTable("", id="nw-peers")

    def on_mount(self) -> None:
        self.refresh_data()
        self.set_interval(2.0, self.refresh_data)

    def refresh_data(self) -> None:
        st = U.read_p2p_status(self.config)
        self.query_one(P2PStatusPanel).show(st)
        self.query_one(DHTPanel).show(st)
        self.query_one(BandwidthPanel).show(st)
        self.query_one(PeerTable).show(st)
