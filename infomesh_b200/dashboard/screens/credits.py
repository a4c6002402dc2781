"""Credits tab: balance, tier, allowance state and the earnings breakdown; posts ``CreditEarned`` when the balance grows
(reference infomesh/dashboard/screens/credits.py:30-386)."""
from __future__ import annotations

from textual.app import ComposeResult
from textual.containers import Vertical
from textual.message import Message
from textual.widgets import Static

from infomesh_b200.dashboard import utils as U
from infomesh_b200.dashboard.widgets import BarChart


class CreditsPane(Vertical):
    class CreditEarned(Message):
        def __init__(self, amount: float):
            super().__init__()
            self.amount = amount

    def __init__(self, config, **kw):
        super().__init__(**kw)
        self.config = config
        self._last_earned: float | None = None

    def compose(self) -> ComposeResult:
        yield Static("", id="cd-head")
        yield Static("[bold]Earnings by action[/]")
        yield BarChart("", id="cd-actions")

    def on_mount(self) -> None:
        self.refresh_data()
        self.set_interval(3.0, self.refresh_data)

    def refresh_data(self) -> None:
        path = self.config.node.data_dir / "credits.db"
        head = self.query_one("#cd-head", Static)
        if not path.exists():
            head.update("[dim]No credit history yet — start crawling to earn credits.[/]")
            return
        try:
            from infomesh_b200.credits.ledger import CreditLedger

            led = CreditLedger(path)
            try:
                s, al = led.stats(), led.search_allowance()
                by_action = dict(led.earnings_by_action()) if hasattr(led, "earnings_by_action") else {}
            finally:
                led.close()
        except Exception as exc:  # noqa: BLE001
            head.update(f"[red]ledger unavailable: {exc}[/]")
            return
        extra = (f"  ·  grace {al.grace_remaining_hours:.0f}h left" if al.state.value == "grace" and al.grace_remaining_hours is not None
                 else f"  ·  debt {al.debt_amount:.2f}" if al.state.value == "debt" else "")
        head.update(f"balance [bold green]{s.balance:,.2f}[/]  ·  {U.tier_label(s.tier)}  ·  earned {s.total_earned:,.2f}  spent {s.total_spent:,.2f}\n"
                    f"search cost {al.search_cost:.3f} ({al.state.value}){extra}  ·  contribution score {s.contribution_score:,.2f}")
        self.query_one("#cd-actions", BarChart).set_items(sorted(((str(k), float(v)) for k, v in by_action.items()), key=lambda kv: -kv[1])[:8])
        if self._last_earned is not None and s.total_earned > self._last_earned:
            self.post_message(self.CreditEarned(s.total_earned - self._last_earned))
        self._last_earned = s.total_earned
