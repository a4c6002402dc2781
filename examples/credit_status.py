"""Credits: earn by contributing, spend on searches, inspect the allowance state, build and verify a signed proof."""
import tempfile
from pathlib import Path

from infomesh_b200.credits.ledger import ActionType, CreditLedger
from infomesh_b200.credits.verification import CreditProofBuilder
from infomesh_b200.p2p.keys import KeyPair

with tempfile.TemporaryDirectory() as d:
    kp = KeyPair.generate()
    ledger = CreditLedger(Path(d) / "credits.db")
    for _ in range(5):
        ledger.record_action(ActionType.CRAWL, quantity=1.0, note="example", key_pair=kp)
    ledger.record_action(ActionType.NETWORK_UPTIME, quantity=2.0, key_pair=kp)
    ledger.spend(0.3, reason="search")
    st, al = ledger.stats(), ledger.search_allowance()
    print(f"balance {st.balance:.2f}  earned {st.total_earned:.2f}  spent {st.total_spent:.2f}  tier {st.tier.name}")
    print(f"allowance: {al.state.value}, one search costs {al.search_cost:.3f}")
    proof = CreditProofBuilder(ledger, kp).build_proof(sample_size=3)
    print("proof verifies:", CreditProofBuilder.verify_proof(proof).verified)
    ledger.close()
